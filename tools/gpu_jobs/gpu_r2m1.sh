# round-2 call M1 (1 GPU): the whole GPU suite, smoke, default bench, reference arm, prefill line -- with the code as committed
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/pytest_gpu_m.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_m.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_m.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_m.log
timeout 400 python bench.py > gpurun_out/bench_big_m.json 2> gpurun_out/bench_big_m.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_big_m.json
timeout 400 python bench.py --impl reference > gpurun_out/bench_ref_m.json 2> gpurun_out/bench_ref_m.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/bench_ref_m.json
timeout 200 python bench.py --mode prefill --steps 3 --warmup 3 > gpurun_out/bench_prefill_m.json 2> gpurun_out/bench_prefill_m.err; echo "prefill rc=$?"; cut -c1-200 gpurun_out/bench_prefill_m.json
timeout 300 python tools/batch_decode_bench.py --batches 1,2,4,8,16,32,64 > gpurun_out/batch_decode_m.json 2> gpurun_out/batch_decode_m.err; echo "batch: $(cat gpurun_out/batch_decode_m.json | cut -c1-300)"
timeout 200 python bench.py --mode prefill --tier big --steps 3 --warmup 2 > gpurun_out/bench_prefill_big_1gpu.json 2> gpurun_out/bench_prefill_big_1gpu.err; echo "prefill big 1gpu rc=$?"; cut -c1-200 gpurun_out/bench_prefill_big_1gpu.json
