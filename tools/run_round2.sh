# Round-2 verification on one B200.  Outputs under gpurun_out/$1 (default r2a), summarised into profiles/.
O=gpurun_out/${1:-r2a}; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
timeout 600 python tools/audit_exact.py --out $O/audit_exact.log > $O/audit_stdout.log 2>&1; echo "audit rc=$?"; tail -2 $O/audit_stdout.log
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"; tail -c 600 $O/bench_default.err
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_smoke.py > $O/sanitizer_$tool.log 2>&1; echo "$tool rc=$?"; tail -4 $O/sanitizer_$tool.log
done
