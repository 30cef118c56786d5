# round-2 call N (1 GPU, short): ncu captures of the shipped tcgen05 GEMM (wide, two tiles per CTA; tall) and the launch list of the goldie prefill step
mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm2_kernel -s 50 -c 1 -o gpurun_out/prof_gemm2_wide2 python tools/config5_sweep.py --tiers big --dtypes q4_0 --batches 64 --rows 2048 --out gpurun_out/tmp_sweep.md > gpurun_out/ncu_gemm2_wide2.log 2>&1; echo "ncu wide rc=$?"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_prefill_goldie.csv python bench.py --mode prefill --steps 1 --warmup 1 --no-parity > gpurun_out/ncu_prefill.log 2>&1; echo "ncu list rc=$?"
