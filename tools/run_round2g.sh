# 2-GPU validation: NCCL tests, default bench at N=2 (DP check, all-reduce overlap), ring-depth experiment on GPU 0
O=gpurun_out/${1:-r2g}; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_dist.py tests/test_gpu_runtime.py "tests/test_gpu_parity.py::test_modules_match_reference_golden" -m gpu -q > $O/pytest_2gpu.log 2>&1; echo "pytest rc=$?"; tail -6 $O/pytest_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 > $O/bench_2gpu.json 2> $O/bench_2gpu.err; echo "bench2 rc=$?"; tail -c 600 $O/bench_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > $O/bench_2gpu_ref.json 2> $O/bench_2gpu_ref.err; echo "ref2 rc=$?"
bash tools/run_round2f.sh r2f
