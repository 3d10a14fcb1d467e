O=gpurun_out/${1:-r2d}; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_soft.py -q -x > $O/pytest_soft.log 2>&1; echo "soft rc=$?"; tail -25 $O/pytest_soft.log
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_soft.py > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 $O/pytest_gpu.log
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"; tail -c 1500 $O/bench_default.err
