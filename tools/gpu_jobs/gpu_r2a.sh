# round-2 call A: producer-side fragments vs the round-1 library, math-only / L2-fed forensics, parity tests, smoke, bench both arms
mkdir -p gpurun_out
R1=nanollama_b200/build/variants/lib_r1.so
timeout 900 python tools/decode_ab.py --tier big --layers 10 --variants "NL_LIB=$R1;NL_TILE_IMG=0;NL_TILE_IMG=1;NL_LIB=$R1,NL_TILE_DBG=1;NL_LIB=$R1,NL_TILE_DBG=2;NL_TILE_DBG=1;NL_TILE_DBG=2" > gpurun_out/ab_a.log 2>&1; cat gpurun_out/ab_a.log
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu_a.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu_a.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/bench_big_a.json 2> gpurun_out/bench_big_a.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_big_a.json; tail -3 gpurun_out/bench_big_a.err
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_a.json 2> gpurun_out/bench_ref_a.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_ref_a.json
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tail -2
nproc
