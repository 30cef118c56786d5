for cfg in "1 0" "1 1"; do set -- $cfg; echo "REPS=$1 HOLD=$2"; NL_REPS=$1 NL_HOLD=$2 timeout 200 python bench.py --steps 3 --no-cpu-baseline 2>/dev/null | cut -c1-80; done
NL_REPS=1 NL_HOLD=0 timeout 200 python bench.py --steps 3 --tier mini --no-cpu-baseline 2>/dev/null | cut -c1-80
timeout 600 python -m pytest tests -m gpu -q -x --timeout 120 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
