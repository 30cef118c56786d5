# round-2 call F: sanitizer passes, batched decode through the GEMMs, config-5 sweep, ncu launch list + full capture
mkdir -p gpurun_out
for tool in memcheck initcheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --log-file gpurun_out/sanitizer_$tool.log python tools/sanitizer_target.py > gpurun_out/sanitizer_$tool.out 2>&1; echo "$tool rc=$? $(tail -1 gpurun_out/sanitizer_$tool.out)"; tail -3 gpurun_out/sanitizer_$tool.log
done
NL_TILE_POLL=0 timeout 600 compute-sanitizer --tool memcheck --log-file gpurun_out/sanitizer_memcheck_barrier.log python tools/sanitizer_target.py > gpurun_out/sanitizer_memcheck_barrier.out 2>&1; echo "memcheck barrier rc=$?"; tail -2 gpurun_out/sanitizer_memcheck_barrier.log
for bm in 0 16 8 4; do NL_BATCH_GEMM_MIN=$bm timeout 400 python tools/batch_decode_bench.py > gpurun_out/batch_decode_gemm$bm.json 2> gpurun_out/batch_decode_gemm$bm.err; echo "batch gemm_min=$bm: $(cat gpurun_out/batch_decode_gemm$bm.json | cut -c1-900)"; done
NL_BATCH_GEMM_MIN=2 timeout 600 python -m pytest tests -m gpu -q --timeout 600 -k "batch_forward or batcher" > gpurun_out/pytest_gpu_f.log 2>&1; echo "pytest batch-gemm rc=$?"; tail -4 gpurun_out/pytest_gpu_f.log
timeout 1500 python tools/config5_sweep.py --out gpurun_out/config5_sweep.md > gpurun_out/config5_sweep.log 2>&1; tail -2 gpurun_out/config5_sweep.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"decode_tiled|embed_kernel|argmax|bump_epoch|feed_prompt" -c 300 --csv --log-file gpurun_out/launches_big_decode.csv python bench.py --steps 1 --warmup 1 --tokens-per-step 32 --no-cpu-baseline --no-parity > gpurun_out/ncu_big.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_tiled -s 20 -c 1 -o gpurun_out/prof_tiled_r2 python bench.py --steps 1 --warmup 1 --tokens-per-step 32 --no-cpu-baseline --no-parity > gpurun_out/ncu_tiled.log 2>&1; echo "ncu full rc=$?"
