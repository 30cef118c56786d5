# round-2 call P (2 GPUs, short): tensor-parallel test and TP2 prefill line with the tile-order planes as the default
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x --timeout 250 -k "tensor_parallel" > gpurun_out/pytest_gpu_p.log 2>&1; echo "pytest tp rc=$?"; tail -2 gpurun_out/pytest_gpu_p.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 bench.py --mode prefill --tier big --gpus 2 --steps 3 --warmup 2 > gpurun_out/bench_prefill_big_tp2_p.json 2> gpurun_out/bench_prefill_big_tp2_p.err; echo "prefill tp2 rc=$?"; cut -c1-200 gpurun_out/bench_prefill_big_tp2_p.json
