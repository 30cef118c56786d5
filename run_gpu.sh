set -x
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/pytest_gpu.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_big.csv python bench.py --steps 1 --warmup 1 --tokens-per-step 2 --no-cpu-baseline > gpurun_out/ncu_big.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_big.log
