O=gpurun_out/${1:-r2i}; mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 > $O/bench_8gpu.json 2> $O/bench_8gpu.err; echo "bench8 rc=$?"; tail -c 400 $O/bench_8gpu.err; head -c 600 $O/bench_8gpu.json
