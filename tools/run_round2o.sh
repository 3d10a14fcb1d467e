O=gpurun_out/${1:-r2o}; mkdir -p $O
python tools/timeline.py 400 f32 2>&1 | grep -v -i warn | tail -8
python tools/timeline.py 16384 f32 2>&1 | grep -v -i warn | tail -8
timeout 600 python tools/refine_probe.py > $O/refine_probe.log 2>&1; echo "probe rc=$?"; cat $O/refine_probe.log | cut -c1-330
timeout 300 python tools/fold_probe2.py 2>&1 | grep -v -i warn | tail -16 | cut -c1-250
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest.log
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; python tools/bench_summary.py $O/bench.json 2>/dev/null | head -40
