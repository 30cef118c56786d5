V=nanollama_b200/build/variants
timeout 500 python tools/decode_ab.py --tier big --layers 10 --timeout 100 --variants "NL_LIB=$V/lib_xb1.so;NL_LIB=$V/lib_cur.so;NL_LIB=$V/lib_cur.so,NL_ATT_CHUNK=48;NL_LIB=$V/lib_cur.so,NL_TILE_POLL=0" 2>&1 | tee gpurun_out/ab4.log
NL_LIB=$V/lib_tr.so timeout 200 python tools/decode_ab.py --tier big --layers 10 --timeout 100 --steps 64 --reps 1 --trace gpurun_out/tr --variants "NL_TILE_POLL=1" > gpurun_out/ab_trace.log 2>&1
python tools/trace_summary.py gpurun_out/tr/trace_0.bin > gpurun_out/trace_poll2.md 2>/dev/null; grep "^|\|grid" gpurun_out/trace_poll2.md
python tools/trace_fine.py gpurun_out/tr/trace_0.bin.ck > gpurun_out/trace_fine_poll2.md 2>&1; cat gpurun_out/trace_fine_poll2.md
