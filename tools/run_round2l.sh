O=gpurun_out/${1:-r2l}; mkdir -p $O
timeout 300 python tools/fold_probe2.py > $O/fold_probe2.log 2>&1; echo "probe2 rc=$?"; cat $O/fold_probe2.log | cut -c1-260
timeout 300 python tools/fold_probe.py > $O/fold_probe.log 2>&1; echo "probe rc=$?"; head -12 $O/fold_probe.log
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest.log
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; python tools/bench_summary.py $O/bench.json 2>/dev/null | head -40
