# round-2 call H (2 GPUs, short): lane-per-position attention scores, lean tensor-parallel instantiation, TP parity + TP2 bench
mkdir -p gpurun_out
V=nanollama_b200/build/variants
timeout 200 python tools/decode_ab.py --tier big --layers 10 --timeout 60 --variants "NL_LIB=$V/lib_r1.so;NL_TILE_L2PF=4" > gpurun_out/ab_h.log 2>&1; cat gpurun_out/ab_h.log
timeout 500 python -m pytest tests -m gpu -q -x --timeout 300 -k "tensor_parallel or forward_logits or greedy_stream or long_context or attention_to_the_end or decode_modes" > gpurun_out/pytest_gpu_h.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu_h.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_tp2_h.json 2> gpurun_out/bench_tp2_h.err; echo "bench tp2 rc=$?"; cut -c1-200 gpurun_out/bench_tp2_h.json; tail -2 gpurun_out/bench_tp2_h.err
NL_TILE_NO_SLIM=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 3 --warmup 3 --no-parity > gpurun_out/bench_tp2_noslim_h.json 2> gpurun_out/bench_tp2_noslim_h.err; echo "bench tp2 general rc=$?"; cut -c1-200 gpurun_out/bench_tp2_noslim_h.json
