# round-2 call M2 (2 GPUs): tensor-parallel tests, TP2 decode and prefill lines -- with the code as committed
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --timeout 400 -k "tensor_parallel" > gpurun_out/pytest_gpu_m2.log 2>&1; echo "pytest tp rc=$?"; tail -3 gpurun_out/pytest_gpu_m2.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tp2_m.json 2> gpurun_out/bench_tp2_m.err; echo "bench tp2 rc=$?"; cut -c1-200 gpurun_out/bench_tp2_m.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 bench.py --mode prefill --tier big --gpus 2 --steps 3 --warmup 2 > gpurun_out/bench_prefill_big_tp2.json 2> gpurun_out/bench_prefill_big_tp2.err; echo "prefill tp2 rc=$?"; cut -c1-200 gpurun_out/bench_prefill_big_tp2.json
