# round-2 evidence run: full GPU suite, 1 M-row audit, sanitizers, ncu launch lists and --set full captures
O=gpurun_out/${1:-r2e}; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 $O/pytest_gpu.log
timeout 600 python tools/audit_exact.py --out $O/audit_exact.log > $O/audit_stdout.log 2>&1; echo "audit rc=$?"; tail -1 $O/audit_stdout.log
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --report-api-errors no --print-limit 100 python tools/sanitize_smoke.py > $O/sanitizer_$tool.log 2>&1; echo "$tool rc=$?"; tail -3 $O/sanitizer_$tool.log
done
B="python bench.py --extras none --no-e2e --no-cpu-baseline"
NCU="ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv"
timeout 200 $NCU --log-file $O/launches_tok400.csv $B --steps 2 --warmup 3 > /dev/null 2>&1; echo "ncu tok rc=$?"
timeout 200 $NCU --log-file $O/launches_train512.csv $B --workload train --codes 512 --steps 2 --warmup 3 > /dev/null 2>&1; echo "ncu train rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_tmem -s 3 -c 1 -f -o $O/tc_tmem_k400_full $B --steps 1 --warmup 3 > $O/ncu_full1.log 2>&1; echo "ncu full tmem rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:apply_runs -s 3 -c 1 -f -o $O/apply_runs_k512_full $B --workload train --codes 512 --steps 1 --warmup 3 > $O/ncu_full2.log 2>&1; echo "ncu full apply rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:backward_kernel -s 3 -c 1 -f -o $O/backward_k512_full $B --workload train --codes 512 --steps 1 --warmup 3 > $O/ncu_full3.log 2>&1; echo "ncu full backward rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 2 -c 1 -f -o $O/tc_gemm_full python -c "
import torch, sys; sys.path.insert(0,'.')
import gesture2vec_b200 as g
A=torch.randn(262144,400,device='cuda'); W=torch.randn(400,400,device='cuda')*0.05
for _ in range(4): g.functional.gemm(A,W)
torch.cuda.synchronize()" > $O/ncu_full4.log 2>&1; echo "ncu full gemm rc=$?"
ls -la $O | head -40
