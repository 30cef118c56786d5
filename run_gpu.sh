set -x
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_tiled -s 8 -c 2 -o gpurun_out/prof_tiled_lm python tools/gemv_bench.py --rows 96000 --cols 4096 --iters 10 > gpurun_out/ncu_tiled_lm.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_tiled_lm.log
for s in "4096 4096" "6144 4096" "22016 4096" "4096 11008" "96000 4096" "48000 1536" "32000 768"; do set -- $s; timeout 120 python tools/gemv_bench.py --rows $1 --cols $2 2>&1 | tail -1; done
