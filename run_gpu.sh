set -x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/tp_worker.py > gpurun_out/tp2.log 2>&1; echo "tp2 rc=$?"; grep "\[tp\]\|TP_PARITY\|Error\|error" gpurun_out/tp2.log | tail -12
