# quick re-verification of the final code: smoke, full GPU suite, memcheck of every kernel variant, default bench line
O=gpurun_out/${1:-r2v}; mkdir -p $O
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.log
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_gpu.log
timeout 1200 compute-sanitizer --tool memcheck --report-api-errors no --print-limit 100 python tools/sanitize_smoke.py > $O/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -2 $O/sanitizer_memcheck.log
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; python tools/bench_summary.py $O/bench.json 2>/dev/null | head -9 | cut -c1-330
