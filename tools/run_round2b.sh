# pair-kernel bring-up: tests, variant timings, large-K experiment, small-batch launch list
O=gpurun_out/${1:-r2b}; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_gpu.log
B="python bench.py --extras none --no-e2e --no-cpu-baseline --steps 20"
for v in tmem pair; do
  timeout 120 $B --variant $v > $O/tok400_$v.json 2> $O/tok400_$v.err; echo "tok400 $v rc=$?"
  timeout 120 $B --variant $v --codes 512 > $O/tok512_$v.json 2> $O/tok512_$v.err; echo "tok512 $v rc=$?"
  timeout 120 $B --variant $v --codes 512 --dtype bf16 > $O/tok512bf16_$v.json 2> $O/tok512bf16_$v.err; echo "tok512bf16 $v rc=$?"
done
for v in auto tmem; do
  timeout 120 $B --variant $v --codes 2048 --rows 1048576 > $O/tok2048_$v.json 2> $O/tok2048_$v.err; echo "tok2048 $v rc=$?"
  timeout 120 $B --variant $v --codes 16384 --rows 1048576 --steps 5 > $O/tok16384_$v.json 2> $O/tok16384_$v.err; echo "tok16384 $v rc=$?"
done
python - <<'PY'
import json,glob,os
for f in sorted(glob.glob(os.environ.get('O','gpurun_out/r2b')+'/tok*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(os.path.basename(f), 'ms/step %.4f kernel %.4f frac %.3f sustained %s clocks %s' % (d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['roofline']['frac'], d.get('sustained',{}).get('ms_per_step'), d['clocks']['sm_mhz']))
    except Exception as e:
        print(f, 'ERR', e)
PY
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_n128.csv python tools/latency_probe.py 4 > $O/latency_probe.log 2>&1; echo "ncu n128 rc=$?"
