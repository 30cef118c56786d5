# round-2 call I (2 GPUs, short): attention shape hoisted out of the phase loop, paired polls of the ranks' partials; TP parity + TP2 bench
mkdir -p gpurun_out
V=nanollama_b200/build/variants
timeout 200 python tools/decode_ab.py --tier big --layers 10 --timeout 60 --variants "NL_LIB=$V/lib_r1.so;NL_TILE_L2PF=4" > gpurun_out/ab_i.log 2>&1; cat gpurun_out/ab_i.log
timeout 400 python -m pytest tests -m gpu -q -x --timeout 240 -k "tensor_parallel or forward_logits or greedy_stream or attention_to_the_end or decode_modes" > gpurun_out/pytest_gpu_i.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu_i.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tp2_i.json 2> gpurun_out/bench_tp2_i.err; echo "bench tp2 rc=$?"; cut -c1-200 gpurun_out/bench_tp2_i.json; tail -2 gpurun_out/bench_tp2_i.err
