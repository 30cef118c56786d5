# round-2 call L2 (8 GPUs, short): tensor-parallel one-pass prefill of big at 8 ranks (K = 1376 shards now take the GEMM path)
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --mode prefill --tier big --gpus 8 --steps 3 --warmup 2 > gpurun_out/bench_prefill_big_tp8.json 2> gpurun_out/bench_prefill_big_tp8.err; echo "prefill tp8 rc=$?"; cut -c1-200 gpurun_out/bench_prefill_big_tp8.json; tail -2 gpurun_out/bench_prefill_big_tp8.err
