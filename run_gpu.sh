timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 -k "tensor_core or prefill" > gpurun_out/pytest_prefill.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_prefill.log
timeout 300 python tools/prefill_bench.py --tier goldie --tokens 2047 2>&1 | tail -1
timeout 300 python tools/prefill_bench.py --tier mini --tokens 512 2>&1 | tail -1
