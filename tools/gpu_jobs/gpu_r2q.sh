# round-2 call Q (1 GPU, one minute): one-pass prefill of big on one GPU with the tile-order planes
mkdir -p gpurun_out
timeout 100 python bench.py --mode prefill --tier big --steps 3 --warmup 2 > gpurun_out/bench_prefill_big_1gpu_q.json 2> gpurun_out/bench_prefill_big_1gpu_q.err; echo "prefill big 1gpu rc=$?"; cut -c1-200 gpurun_out/bench_prefill_big_1gpu_q.json
