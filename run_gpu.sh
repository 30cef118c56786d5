# Round-end validation on one B200 (what `gpurun -- 'bash run_gpu.sh'` ran for the records under profiles/):
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_big.json 2> gpurun_out/bench_big.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/bench_big.json
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 2>/dev/null | tail -1 | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_tiled -s 20 -c 1 -o gpurun_out/prof_tiled python bench.py --steps 1 --warmup 1 --tokens-per-step 32 --no-cpu-baseline > gpurun_out/ncu_tiled.log 2>&1; echo "ncu full rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"decode_tiled|embed_kernel|argmax|bump_epoch|feed_prompt" -c 300 --csv --log-file gpurun_out/launches_big_decode.csv python bench.py --steps 1 --warmup 1 --tokens-per-step 32 --no-cpu-baseline > gpurun_out/ncu_big.log 2>&1; echo "ncu list rc=$?"
# A/B of kernel variants (10-layer big): python tools/decode_ab.py --tier big --layers 10 --variants "NL_TILE_POLL=0;NL_TILE_POLL=1"
# other records: tools/sample_bench.py, tools/batch_decode_bench.py, tools/gemv_sweep.py, tools/prefill_bench.py, tools/decode_ctx_bench.py
