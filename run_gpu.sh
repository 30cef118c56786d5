timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "matmul or forward or greedy or wide_tier or long_context or reset or tied or gamma or generate" > gpurun_out/pytest_quick.log 2>&1; echo "quick rc=$?"; tail -6 gpurun_out/pytest_quick.log
timeout 120 python tools/gemv_bench.py --rows 96000 --cols 4096 --dtype q8_0 2>&1 | tail -1
timeout 120 python tools/gemv_bench.py --rows 96000 --cols 4096 --dtype q4_0 2>&1 | tail -1
timeout 300 python bench.py --tier mini --dtype q8_0 --steps 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-160
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'])"
