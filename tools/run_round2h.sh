O=gpurun_out/${1:-r2h}; mkdir -p $O
timeout 300 python tools/fold_probe.py > $O/fold_probe.log 2>&1; echo "probe rc=$?"; cat $O/fold_probe.log | tail -12
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_runtime.py tests/test_kmeans_tokenizer.py -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest.log
