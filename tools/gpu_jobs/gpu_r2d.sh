# round-2 call D: lean finishing warp + warp-per-position attention scores + 112-register cap, forensics modes, traces, parity, bench lines
mkdir -p gpurun_out
V=nanollama_b200/build/variants
timeout 900 python tools/decode_ab.py --tier big --layers 10 --variants "NL_LIB=$V/lib_r1.so;NL_LIB=$V/lib_n1.so;NL_LIB=$V/lib_n1lb.so;NL_LIB=$V/lib_n1spin.so;NL_LIB=$V/lib_n1.so,NL_TILE_DBG=1;NL_LIB=$V/lib_n1.so,NL_TILE_DBG=4;NL_LIB=$V/lib_n1skip.so,NL_TILE_DBG=1;NL_LIB=$V/lib_n1skip.so,NL_TILE_DBG=4;NL_LIB=$V/lib_n1.so,NL_TILE_INFLIGHT=2" > gpurun_out/ab_d.log 2>&1; cat gpurun_out/ab_d.log
timeout 300 python tools/decode_ab.py --tier big --layers 10 --trace gpurun_out/trace_d --variants "NL_LIB=$V/lib_n1tr.so" > gpurun_out/ab_d_trace.log 2>&1; cat gpurun_out/ab_d_trace.log
python tools/trace_summary.py gpurun_out/trace_d/trace_0.bin > gpurun_out/trace_d/summary_0.md 2>&1; python tools/trace_fine.py gpurun_out/trace_d/trace_0.bin.ck > gpurun_out/trace_d/fine_0.md 2>&1; rm -f gpurun_out/trace_d/*.bin gpurun_out/trace_d/*.ck
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -k "forward or greedy or decode_modes or wide_tier or long_context or full_depth or attention_to_the_end or bias or gamma or reset" > gpurun_out/pytest_gpu_d.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu_d.log
timeout 600 python bench.py > gpurun_out/bench_big_d.json 2> gpurun_out/bench_big_d.err; echo "bench rc=$?"; cut -c1-250 gpurun_out/bench_big_d.json; tail -3 gpurun_out/bench_big_d.err
for t in goldie mini; do timeout 300 python bench.py --tier $t --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${t}_d.json 2> gpurun_out/bench_${t}_d.err; cut -c1-200 gpurun_out/bench_${t}_d.json; done
