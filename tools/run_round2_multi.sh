# multi-GPU bench lines (weak scaling): torchrun, one rank per GPU
N=${1:-2}; O=gpurun_out/r2z_${N}gpu; mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
python tools/bench_summary.py $O/bench.json 2>/dev/null | head -9 | cut -c1-400; nvidia-smi topo -m 2>/dev/null | head -14 > $O/topo.txt
grep -o '"dp_check": {[^}]*}' $O/bench.json | cut -c1-400
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q 2>&1 | tail -2
