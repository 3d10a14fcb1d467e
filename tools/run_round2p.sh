O=gpurun_out/${1:-r2p}; mkdir -p $O
B="python bench.py --extras none --no-e2e --no-cpu-baseline"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:apply_runs -s 3 -c 1 -f -o $O/apply_runs_k512_full $B --workload train --codes 512 --steps 1 --warmup 3 > $O/ncu_full2.log 2>&1; echo "ncu full apply rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:backward_flat -s 3 -c 1 -f -o $O/backward_k512_full $B --workload train --codes 512 --steps 1 --warmup 3 > $O/ncu_full3.log 2>&1; echo "ncu full backward rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rerank_kernel -s 3 -c 1 -f -o $O/rerank_k400_full $B --steps 1 --warmup 3 > $O/ncu_full5.log 2>&1; echo "ncu full rerank rc=$?"
ls -la $O
