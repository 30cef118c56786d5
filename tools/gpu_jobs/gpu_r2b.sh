# round-2 call B: HMMA microbenchmark, streaming-loop variants, forensics modes, fine traces, parity tests of the merged kernel
mkdir -p gpurun_out
V=nanollama_b200/build/variants
nanollama_b200/build/bin/ubench_mma > gpurun_out/ubench_mma.jsonl 2>&1; cat gpurun_out/ubench_mma.jsonl
timeout 1200 python tools/decode_ab.py --tier big --layers 10 --variants "NL_LIB=$V/lib_r1.so;NL_TILE_IMG=1;NL_TILE_IMG=0;NL_LIB=$V/lib_x64.so,NL_TILE_IMG=0;NL_LIB=$V/lib_acc4.so,NL_TILE_IMG=0;NL_LIB=$V/lib_acc8.so,NL_TILE_IMG=0;NL_LIB=$V/lib_nv.so,NL_TILE_IMG=0;NL_LIB=$V/lib_acc4nv.so,NL_TILE_IMG=0;NL_TILE_IMG=0,NL_TILE_DBG=1;NL_TILE_IMG=0,NL_TILE_DBG=3;NL_LIB=$V/lib_skipmath.so,NL_TILE_IMG=0,NL_TILE_DBG=1;NL_LIB=$V/lib_skipmath.so,NL_TILE_IMG=0,NL_TILE_DBG=3;NL_LIB=$V/lib_acc4.so,NL_TILE_IMG=1" > gpurun_out/ab_b.log 2>&1; cat gpurun_out/ab_b.log
timeout 400 python tools/decode_ab.py --tier big --layers 10 --trace gpurun_out/trace_b --variants "NL_LIB=$V/lib_tr.so,NL_TILE_IMG=0;NL_LIB=$V/lib_tr.so,NL_TILE_IMG=1" > gpurun_out/ab_b_trace.log 2>&1; cat gpurun_out/ab_b_trace.log
for i in 0 1; do python tools/trace_summary.py gpurun_out/trace_b/trace_$i.bin > gpurun_out/trace_b/summary_$i.md 2>&1; python tools/trace_fine.py gpurun_out/trace_b/trace_$i.bin.ck > gpurun_out/trace_b/fine_$i.md 2>&1; done
rm -f gpurun_out/trace_b/*.bin gpurun_out/trace_b/*.ck
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu_b.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_b.log
