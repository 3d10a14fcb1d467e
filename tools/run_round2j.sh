O=gpurun_out/${1:-r2j}; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 $O/pytest_gpu.log
timeout 300 python tools/fold_probe.py > $O/fold_probe.log 2>&1; echo "probe rc=$?"; tail -9 $O/fold_probe.log
timeout 300 python bench.py --workload train --codes 512 --no-e2e --no-cpu-baseline --steps 10 > $O/train512.json 2> $O/train512.err; echo "train rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2j/train512.json').read().strip().splitlines()[-1])
print('train512 ms/step', d['ms_per_step'], 'launches/step', d['gpu_launches']/d['steps'], 'clk', d['clocks']['sm_mhz'])
PY
