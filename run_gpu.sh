set -x
timeout 600 python bench.py --steps 5 > gpurun_out/bench_big.json 2> gpurun_out/bench_big.err; echo "bench big rc=$?"; cat gpurun_out/bench_big.json; tail -3 gpurun_out/bench_big.err
timeout 300 python bench.py --impl reference --steps 2 > gpurun_out/bench_big_ref.json 2> gpurun_out/bench_big_ref.err; echo "bench ref rc=$?"; cat gpurun_out/bench_big_ref.json | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 500 --csv --log-file gpurun_out/launches_big.csv python bench.py --steps 1 --warmup 1 --tokens-per-step 2 --no-cpu-baseline > gpurun_out/ncu_big.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemv_stream -s 196 -c 6 -o gpurun_out/prof_stream python bench.py --steps 1 --warmup 1 --tokens-per-step 2 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
