# round-2 call C (2 GPUs): polled tensor-parallel exchange + tensor-parallel prefill parity, TP2 bench, single-GPU regression check,
# tensor-core prefill attention, HMMA microbenchmark
mkdir -p gpurun_out
V=nanollama_b200/build/variants
nanollama_b200/build/bin/ubench_mma > gpurun_out/ubench_mma.jsonl 2>&1; cat gpurun_out/ubench_mma.jsonl
timeout 300 python tools/decode_ab.py --tier big --layers 10 --variants "NL_LIB=$V/lib_r1.so;NL_TILE_IMG=0;NL_TILE_IMG=1" > gpurun_out/ab_c.log 2>&1; cat gpurun_out/ab_c.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -k "tensor_parallel or prefill or batcher or bias or reset_and_replay" > gpurun_out/pytest_gpu_c.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/pytest_gpu_c.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_tp2_c.json 2> gpurun_out/bench_tp2_c.err; echo "bench tp2 rc=$?"; cut -c1-400 gpurun_out/bench_tp2_c.json; tail -5 gpurun_out/bench_tp2_c.err
NL_TILE_POLL=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --no-parity > gpurun_out/bench_tp2_barrier_c.json 2> gpurun_out/bench_tp2_barrier_c.err; echo "bench tp2 barrier rc=$?"; cut -c1-200 gpurun_out/bench_tp2_barrier_c.json
timeout 600 python bench.py --mode prefill --steps 3 --warmup 1 > gpurun_out/bench_prefill_c.json 2> gpurun_out/bench_prefill_c.err; echo "prefill rc=$?"; cut -c1-600 gpurun_out/bench_prefill_c.json; tail -3 gpurun_out/bench_prefill_c.err
NL_PREFILL_ATTN_CC=1 timeout 600 python bench.py --mode prefill --steps 3 --warmup 1 > gpurun_out/bench_prefill_cc_c.json 2> gpurun_out/bench_prefill_cc_c.err; echo "prefill cc rc=$?"; cut -c1-300 gpurun_out/bench_prefill_cc_c.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --mode prefill --tier big --steps 2 --warmup 1 > gpurun_out/bench_prefill_tp2_c.json 2> gpurun_out/bench_prefill_tp2_c.err; echo "prefill tp2 rc=$?"; cut -c1-400 gpurun_out/bench_prefill_tp2_c.json; tail -3 gpurun_out/bench_prefill_tp2_c.err
