# round-2 call G (short, tight timeouts): lean kernel with the new attention scores, the general instantiation fixed, parity subset, bench
mkdir -p gpurun_out
V=nanollama_b200/build/variants
timeout 420 python tools/decode_ab.py --tier big --layers 10 --timeout 60 --variants "NL_LIB=$V/lib_r1.so;NL_LIB=$V/lib_n3.so;NL_LIB=$V/lib_n3.so,NL_TILE_L2PF=0;NL_LIB=$V/lib_n3.so,NL_TILE_NO_SLIM=1;NL_LIB=$V/lib_n3.so,NL_TILE_DBG=1;NL_LIB=$V/lib_n3.so,NL_TILE_DBG=4" > gpurun_out/ab_g.log 2>&1; cat gpurun_out/ab_g.log
timeout 120 python tools/decode_ab.py --tier big --layers 10 --timeout 60 --trace gpurun_out/trace_g --variants "NL_LIB=$V/lib_n3tr.so" > gpurun_out/ab_g_trace.log 2>&1; cat gpurun_out/ab_g_trace.log
python tools/trace_summary.py gpurun_out/trace_g/trace_0.bin > gpurun_out/trace_g/summary_0.md 2>&1; python tools/trace_fine.py gpurun_out/trace_g/trace_0.bin.ck > gpurun_out/trace_g/fine_0.md 2>&1; rm -f gpurun_out/trace_g/*.bin gpurun_out/trace_g/*.ck
timeout 600 python -m pytest tests -m gpu -q -x --timeout 240 -k "forward or greedy or decode_modes or wide_tier or long_context or attention_to_the_end or bias or gamma or reset" > gpurun_out/pytest_gpu_g.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu_g.log
timeout 300 python bench.py > gpurun_out/bench_big_g.json 2> gpurun_out/bench_big_g.err; echo "bench rc=$?"; cut -c1-250 gpurun_out/bench_big_g.json; tail -3 gpurun_out/bench_big_g.err
for t in goldie mini; do timeout 120 python bench.py --tier $t --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${t}_g.json 2> gpurun_out/bench_${t}_g.err; cut -c1-200 gpurun_out/bench_${t}_g.json; done
