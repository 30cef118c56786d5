timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 200 -k "sampling" 2>&1 | tail -8
timeout 200 python tools/sample_bench.py 2>&1 | tail -1 | tee gpurun_out/sample_bench.json
