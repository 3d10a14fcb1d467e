O=gpurun_out/${1:-r2c}; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_gemm.py -q -x > $O/pytest_gemm.log 2>&1; echo "gemm rc=$?"; tail -15 $O/pytest_gemm.log
timeout 600 python -m pytest tests -m gpu -q --deselect tests/test_gpu_gemm.py > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_gpu.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_n128.csv python tools/latency_probe.py 4 > $O/latency_probe.log 2>&1; echo "ncu n128 rc=$?"
B="python bench.py --extras none --no-e2e --no-cpu-baseline --steps 10"
timeout 120 $B --codes 2048 --rows 1048576 > $O/tok2048_auto.json 2> $O/tok2048_auto.err; echo "tok2048 rc=$?"
timeout 120 $B --codes 2048 --rows 1048576 --dtype bf16 > $O/tok2048bf16_auto.json 2> $O/tok2048bf16.err; echo "tok2048bf16 rc=$?"
