# round-2 evidence run: full GPU suite, 1 M-row audit, sanitizers, ncu launch lists and --set full captures, bench
O=gpurun_out/${1:-r2z}; mkdir -p $O
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu.log
timeout 600 python tools/audit_exact.py --out $O/audit_exact.log > $O/audit_stdout.log 2>&1; echo "audit rc=$?"; tail -1 $O/audit_stdout.log
for tool in memcheck racecheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool --report-api-errors no --print-limit 100 python tools/sanitize_smoke.py > $O/sanitizer_$tool.log 2>&1; echo "$tool rc=$?"; tail -3 $O/sanitizer_$tool.log
done
B="python bench.py --extras none --no-e2e --no-cpu-baseline"
NCU="ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv"
timeout 200 $NCU --log-file $O/launches_tok400.csv $B --steps 2 --warmup 3 > /dev/null 2>&1; echo "ncu tok rc=$?"
timeout 200 $NCU --log-file $O/launches_train512.csv $B --workload train --codes 512 --steps 2 --warmup 3 > /dev/null 2>&1; echo "ncu train rc=$?"
FULL="ncu --set full --clock-control none --import-source on -c 1 -f"
timeout 300 $FULL -k regex:tc_tmem -s 3 -o $O/tc_tmem_k400_full $B --steps 1 --warmup 3 > $O/ncu_full1.log 2>&1; echo "ncu full tmem rc=$?"
timeout 300 $FULL -k regex:rerank_kernel -s 3 -o $O/rerank_k400_full $B --steps 1 --warmup 3 > $O/ncu_full2.log 2>&1; echo "ncu full rerank rc=$?"
timeout 300 $FULL -k regex:apply_runs -s 3 -o $O/apply_runs_k512_full $B --workload train --codes 512 --steps 1 --warmup 3 > $O/ncu_full3.log 2>&1; echo "ncu full apply rc=$?"
timeout 300 $FULL -k regex:backward_flat -s 3 -o $O/backward_k512_full $B --workload train --codes 512 --steps 1 --warmup 3 > $O/ncu_full4.log 2>&1; echo "ncu full backward rc=$?"
timeout 300 $FULL -k regex:refine_rows -s 3 -o $O/refine_rows_k16384_full $B --codes 16384 --steps 1 --warmup 3 > $O/ncu_full5.log 2>&1; echo "ncu full refine rows rc=$?"
timeout 300 $FULL -k regex:tc_gemm -s 3 -o $O/refine_gemm_k16384_full $B --codes 16384 --steps 1 --warmup 3 > $O/ncu_full6.log 2>&1; echo "ncu full refine gemm rc=$?"
timeout 300 $FULL -k regex:tc_gemm -s 2 -o $O/tc_gemm_full python -c "
import torch, sys; sys.path.insert(0,'.')
import gesture2vec_b200 as g
A=torch.randn(262144,400,device='cuda'); W=torch.randn(400,400,device='cuda')*0.05
for _ in range(4): g.functional.gemm(A,W)
torch.cuda.synchronize()" > $O/ncu_full7.log 2>&1; echo "ncu full gemm rc=$?"
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; python tools/bench_summary.py $O/bench.json 2>/dev/null | head -12
timeout 600 python bench.py --workload kmeans --codes 300 --extras none > $O/bench_kmeans300.json 2> $O/bench_kmeans.err; echo "bench kmeans rc=$?"; python tools/bench_summary.py $O/bench_kmeans300.json 2>/dev/null | head -4 | cut -c1-300
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "bench ref rc=$?"; cut -c1-400 $O/bench_ref.json
ls -la $O | head -50
