mkdir -p gpurun_out/r1v
O=gpurun_out/r1v
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
timeout 300 python bench.py > $O/bench_tok400.json 2> $O/bench_tok400.err; echo "b1 rc=$?"
timeout 300 python bench.py --workload train --codes 512 --no-cpu-baseline > $O/bench_train512.json 2> $O/bench_train512.err; echo "b2 rc=$?"
timeout 300 python bench.py --codes 16384 --rows 262144 --no-cpu-baseline > $O/bench_tok16384.json 2> $O/bench_tok16384.err; echo "b3 rc=$?"
timeout 300 python bench.py --codes 512 --dtype bf16 --no-cpu-baseline > $O/bench_tok512_bf16.json 2> $O/bench_tok512_bf16.err; echo "b4 rc=$?"
timeout 200 python bench.py --impl reference --steps 5 --warmup 2 > $O/bench_ref.json 2> $O/bench_ref.err; echo "b5 rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_tok400.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_b.log 2>&1; echo "ncu1 rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_ -s 3 -c 1 -f -o $O/tc_k400_full python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_full.log 2>&1; echo "ncu2 rc=$?"
cat $O/bench_tok400.json | cut -c1-1500
