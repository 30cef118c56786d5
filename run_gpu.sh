set -x
timeout 600 python -m pytest tests -m gpu -q -x --timeout 120 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
NL_TRACE=gpurun_out/trace_big.bin timeout 300 python bench.py --steps 1 --warmup 1 --tokens-per-step 8 --no-cpu-baseline > gpurun_out/bench_big_tr.json 2> gpurun_out/bench_big.err; echo "rc=$?"; cut -c1-160 gpurun_out/bench_big_tr.json; tail -3 gpurun_out/bench_big.err
timeout 300 python bench.py --steps 3 --no-cpu-baseline > gpurun_out/bench_big.json 2> gpurun_out/bench_big.err; echo "bench big rc=$?"; cat gpurun_out/bench_big.json | cut -c1-200
timeout 300 python bench.py --tier mini --steps 3 --no-cpu-baseline > gpurun_out/bench_mini.json 2> gpurun_out/bench_mini.err; echo "bench mini rc=$?"; cat gpurun_out/bench_mini.json | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_mega -s 20 -c 1 -o gpurun_out/prof_mega3 python bench.py --steps 1 --warmup 1 --tokens-per-step 4 --no-cpu-baseline > gpurun_out/ncu_mega.log 2>&1; echo "ncu rc=$?"
