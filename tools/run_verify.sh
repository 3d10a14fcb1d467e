# Round verification on one B200: GPU tests, the bench lines of BASELINE.json's configs, ncu launch lists and one
# --set full capture of the dominant kernel.  Outputs under gpurun_out/r1f (copied / summarised into profiles/).
O=gpurun_out/r1f; mkdir -p $O
timeout 400 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_gpu.log
timeout 200 python bench.py > $O/bench_tok400.json 2> $O/bench_tok400.err; echo "tok400 rc=$?"
timeout 200 python bench.py --workload train --codes 512 --no-cpu-baseline > $O/bench_train512.json 2> $O/bench_train512.err; echo "train512 rc=$?"
timeout 200 python bench.py --codes 16384 --rows 262144 --no-cpu-baseline > $O/bench_tok16384.json 2> $O/bench_tok16384.err; echo "tok16384 rc=$?"
timeout 200 python bench.py --codes 512 --dtype bf16 --no-cpu-baseline > $O/bench_tok512_bf16.json 2> $O/bench_tok512_bf16.err; echo "bf16 rc=$?"
timeout 200 python bench.py --workload kmeans --codes 300 --cpu-seconds 6 > $O/bench_kmeans300.json 2> $O/bench_kmeans300.err; echo "kmeans rc=$?"
timeout 200 python bench.py --impl reference --steps 5 --warmup 2 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_tok400.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1; echo "ncu1 rc=$?"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_train512.csv python bench.py --workload train --codes 512 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1; echo "ncu2 rc=$?"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_tok16384.csv python bench.py --codes 16384 --rows 262144 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1; echo "ncu3 rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:tc_tmem -s 3 -c 1 -f -o $O/tc_tmem_k400_full python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_full.log 2>&1; echo "ncu4 rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:tc_search -s 3 -c 1 -f -o $O/tc_search_k16384_full python bench.py --codes 16384 --rows 262144 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_full2.log 2>&1; echo "ncu5 rc=$?"
