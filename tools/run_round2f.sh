# small experiments: row-ring depth sensitivity of tc_tmem_kernel
O=gpurun_out/${1:-r2f}; mkdir -p $O
B="python bench.py --extras none --no-e2e --no-cpu-baseline --steps 20"
for z in 3 4 6 10; do
  G2V_TC_ZSLOTS=$z timeout 120 $B > $O/tok400_z$z.json 2> $O/tok400_z$z.err; echo "z$z rc=$?"
done
for b in 2 3 5 8; do
  G2V_TC_BSTAGES=$b timeout 120 $B > $O/tok400_b$b.json 2> $O/tok400_b$b.err; echo "b$b rc=$?"
done
python - <<'PY'
import json,glob,os
for f in sorted(glob.glob('gpurun_out/r2f/tok*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(os.path.basename(f), 'ms/step %.4f kernel %.4f clocks %s' % (d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['clocks']['sm_mhz']))
    except Exception as e:
        print(f, 'ERR', e)
PY
