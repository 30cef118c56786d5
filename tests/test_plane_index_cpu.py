"""Host check of plane_index (nanollama_b200/csrc/nl_common.cuh): the order in which the producers of a GEMM input write its bf16
planes so that nl_gemm2.cuh can fetch a 128-row x 32-k tile with one bulk copy.  The function is __host__ __device__: it is compiled
for the host here with nvcc and compared with the layout the kernel's descriptors assume (K-major, no swizzle: core matrices of
8 rows x 16 bytes, 128 bytes apart along the rows, 2048 bytes apart along K)."""
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r"""
#include <cstdio>
#include <cstdlib>
#include "nl_common.cuh"
int main(int argc, char **argv) {
    const int T = atoi(argv[1]), K = atoi(argv[2]), tiled = atoi(argv[3]);
    for (int r = 0; r < T; r++)
        for (int k = 0; k < K; k++) { const unsigned long long i = nl::plane_index(r, k, K, tiled); fwrite(&i, 8, 1, stdout); }
    return 0;
}
"""


@pytest.fixture(scope="module")
def plane_index_bin(tmp_path_factory):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not found")
    d = tmp_path_factory.mktemp("plane_index")
    src = d / "pi.cu"
    src.write_text(SRC)
    exe = d / "pi"
    subprocess.run([nvcc, "-std=c++17", "-O1", "-I", os.path.join(ROOT, "nanollama_b200", "csrc"), str(src), "-o", str(exe)], check=True, capture_output=True)
    return str(exe)


@pytest.mark.parametrize("T,K", [(300, 1376), (128, 64), (17, 4096), (2047, 96)])
def test_tile_order_is_a_bijection_onto_whole_tiles_in_umma_order(plane_index_bin, T, K):
    out = subprocess.run([plane_index_bin, str(T), str(K), "1"], check=True, capture_output=True).stdout
    idx = np.frombuffer(out, dtype=np.uint64).reshape(T, K).astype(np.int64)
    assert np.unique(idx).size == T * K                       # no two elements share a slot
    assert idx.max() < ((T + 127) // 128 * 128) * K           # inside the planes as allocated (whole 128-row tiles)
    r = np.arange(T)[:, None]
    k = np.arange(K)[None, :]
    tile = (r >> 7) * (K >> 5) + (k >> 5)                     # [row / 128][k / 32] tiles of 4096 elements = 8 KB
    inner_bytes = (idx - tile * 4096) * 2
    exp = ((k & 31) >> 3) * 2048 + ((r & 127) >> 3) * 128 + (r & 7) * 16 + (k & 7) * 2   # what the cp.async path of the kernel writes
    assert np.array_equal(inner_bytes, exp)


def test_row_major_when_not_tiled(plane_index_bin):
    out = subprocess.run([plane_index_bin, "5", "64", "0"], check=True, capture_output=True).stdout
    assert np.array_equal(np.frombuffer(out, dtype=np.uint64), np.arange(5 * 64, dtype=np.uint64))
