O=gpurun_out/${1:-r2m}; mkdir -p $O
timeout 600 python tools/refine_probe.py > $O/refine_probe.log 2>&1; echo "probe rc=$?"; cat $O/refine_probe.log | cut -c1-330
timeout 900 python -m pytest tests/test_gpu_audit.py tests/test_gpu_parity.py -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $O/pytest.log
timeout 300 python tools/fold_probe2.py > $O/fold_probe2.log 2>&1; echo "probe2 rc=$?"; cat $O/fold_probe2.log | cut -c1-260
