timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "wide_tier or forward_logits or greedy_stream" > gpurun_out/pytest_quick.log 2>&1; echo "quick rc=$?"; tail -3 gpurun_out/pytest_quick.log
timeout 120 python tools/gemv_bench.py --rows 96000 --cols 4096 2>&1 | tail -1
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'])"
