echo "--- default descriptor convention"; timeout 120 python tools/gemm_check.py 2>&1 | tail -14
echo "--- swapped LBO/SBO"; NL_GEMM_SWAP=1 timeout 120 python tools/gemm_check.py 2>&1 | head -4
