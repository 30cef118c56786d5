V=nanollama_b200/build/variants
timeout 400 python tools/decode_ab.py --tier big --layers 10 --timeout 100 --variants "NL_LIB=$V/lib_cur.so;NL_LIB=$V/lib_xb2.so;NL_LIB=$V/lib_cur.so" 2>&1 | tee gpurun_out/ab7.log
for L in cur xb2; do for S in "96000 4096" "4096 11008" "22016 4096"; do set -- $S; NL_LIB=$V/lib_$L.so timeout 100 python tools/gemv_bench.py --rows $1 --cols $2 2>&1 | tail -1 | sed "s/^/$L /"; done; done | tee gpurun_out/gemv_ab.log
