# round-2 call K3 (1 GPU): tcgen05 GEMM with 16 producer warps (two threads per weight row), one-wave split K
mkdir -p gpurun_out
timeout 300 python tools/gemm_check.py > gpurun_out/gemm_check_k3.log 2>&1; echo "gemm_check rc=$?"; awk '{print $1,$2,$3,$4,$6}' gpurun_out/gemm_check_k3.log | tr '\n' ';'; echo
timeout 600 python -m pytest tests -m gpu -q -x --timeout 240 -k "prefill or batch or matrix or gemm or bias or mixed" > gpurun_out/pytest_gpu_k3.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_k3.log
timeout 300 python tools/config5_sweep.py --tiers goldie,big --dtypes q4_0 --batches 16,64,128 --rows 512,2048 --out gpurun_out/gemm2c_sweep.md > /dev/null 2>&1; echo "sweep v2c rc=$?"; grep -v "^$" gpurun_out/gemm2c_sweep.md | tail -50 | cut -c1-120
timeout 200 python tools/config5_sweep.py --tiers big --dtypes q8_0,f16 --batches 64 --rows 2048 --out gpurun_out/gemm2c_sweep_q8_f16.md > /dev/null 2>&1; grep "B=64\|M=2048" gpurun_out/gemm2c_sweep_q8_f16.md | cut -c1-120
timeout 200 python bench.py --mode prefill --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_prefill_k3.json 2> gpurun_out/bench_prefill_k3.err; echo "prefill v2c rc=$?"; cut -c1-250 gpurun_out/bench_prefill_k3.json
timeout 300 python tools/batch_decode_bench.py --batches 1,2,4,8,16,32,64 > gpurun_out/batch_decode_k3.json 2> gpurun_out/batch_decode_k3.err; echo "batch: $(cat gpurun_out/batch_decode_k3.json | cut -c1-1200)"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm2_kernel -s 15 -c 1 -o gpurun_out/prof_gemm2_wide python tools/config5_sweep.py --tiers big --dtypes q4_0 --batches 64 --rows 2048 --out gpurun_out/tmp_sweep.md > gpurun_out/ncu_gemm2.log 2>&1; echo "ncu rc=$?"
