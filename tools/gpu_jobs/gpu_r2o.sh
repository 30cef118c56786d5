# round-2 call O (1 GPU, short): activation planes in tile order + bulk copies by the issuing thread (NL_GEMM_ATILE=1), against the default
mkdir -p gpurun_out
export NL_GEMM_ATILE=1
timeout 150 python tools/gemm_check.py > gpurun_out/gemm_check_o.log 2>&1; echo "gemm_check atile rc=$?"; awk '{print $1,$2,$3,$4,$6}' gpurun_out/gemm_check_o.log | tr '\n' ';'; echo
timeout 300 python -m pytest tests -m gpu -q -x --timeout 200 -k "prefill or batch or matrix or many_rows or bias or mixed" > gpurun_out/pytest_gpu_o.log 2>&1; echo "pytest atile rc=$?"; tail -3 gpurun_out/pytest_gpu_o.log
timeout 100 python bench.py --mode prefill --steps 3 --warmup 3 > gpurun_out/bench_prefill_o.json 2> gpurun_out/bench_prefill_o.err; echo "prefill atile rc=$?"; cut -c1-200 gpurun_out/bench_prefill_o.json
timeout 100 python tools/config5_sweep.py --tiers big --dtypes q4_0 --batches 64 --rows 512,2048 --out gpurun_out/gemm2_atile_sweep.md > /dev/null 2>&1; grep "B=64\|M=" gpurun_out/gemm2_atile_sweep.md | cut -c1-120
unset NL_GEMM_ATILE
timeout 100 python -m pytest tests -m gpu -q -x --timeout 100 -k "prefill_one_pass or many_rows" > gpurun_out/pytest_gpu_o2.log 2>&1; echo "pytest default rc=$?"; tail -2 gpurun_out/pytest_gpu_o2.log
