O=gpurun_out/${1:-r2k}; mkdir -p $O
timeout 300 python tools/fold_probe.py > $O/fold_probe.log 2>&1; echo "probe rc=$?"; head -12 $O/fold_probe.log; grep -A30 "cumulative" $O/fold_probe.log | cut -c1-150 | head -40
timeout 600 python -m pytest tests/test_gpu_runtime.py tests/test_gpu_parity.py -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest.log
