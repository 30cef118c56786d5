timeout 45 python tools/batch_decode_bench.py --steps 16 2>&1 | tail -1 | tee gpurun_out/batch_decode.json
