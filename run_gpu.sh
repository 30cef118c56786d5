timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 200 -k "sampling" 2>&1 | tail -15
