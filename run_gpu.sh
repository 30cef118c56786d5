timeout 600 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 700 python tools/decode_ab.py --tier big --layers 10 --timeout 100 --variants "NL_TILE_POLL=0;NL_TILE_POLL=1;NL_TILE_POLL=1,NL_TILE_POLL_NS=100;NL_TILE_POLL=1,NL_ATT_CHUNK=48;NL_TILE_POLL=1,NL_ATT_CHUNK=32;NL_TILE_POLL=0,NL_ATT_CHUNK=32" 2>&1 | tee gpurun_out/ab1.log
timeout 300 python tools/decode_ab.py --tier big --layers 10 --timeout 100 --steps 64 --reps 1 --trace gpurun_out/tr --variants "NL_TILE_POLL=0;NL_TILE_POLL=1" > gpurun_out/ab_trace.log 2>&1
for i in 0 1; do python tools/trace_summary.py gpurun_out/tr/trace_$i.bin > gpurun_out/trace_$i.md 2>&1; done; cat gpurun_out/trace_1.md
