timeout 200 python -m pytest tests/test_gpu_dist.py -q 2>&1 | grep -v "^$" | tail -45
