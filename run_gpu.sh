N=$(nvidia-smi -L | wc -l)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/tp_worker.py > gpurun_out/tp$N.log 2>&1; echo "tp$N parity rc=$?"; grep "\[tp\]\|TP_PARITY\|Error\|error" gpurun_out/tp$N.log | tail -12
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_tp$N.json 2> gpurun_out/bench_tp$N.err; echo "bench tp$N rc=$?"; tail -1 gpurun_out/bench_tp$N.json | cut -c1-200
