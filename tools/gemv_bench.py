#!/usr/bin/env python3
"""Microbench of one dequant-fused GEMV (matmulDispatch, go/model.go:361-386) on L2-cold replicas: GB/s vs the HBM roofline.

    python tools/gemv_bench.py --rows 96000 --cols 4096 --dtype q4_0 [--batch 1] [--iters 40]
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nanollama_b200 import gguf as G
from nanollama_b200 import model as M


def random_blocks(typ, rows, cols, seed=0):
    rng = np.random.default_rng(seed)
    n = rows * cols
    if typ == G.GGML_F16:
        return (rng.standard_normal(n) * 0.02).astype(np.float16).view(np.uint8)
    bs = G.ggml_block_size(typ)
    raw = rng.integers(0, 256, size=(n // 32, bs), dtype=np.uint8)
    raw[:, :2] = (rng.uniform(0.5, 1.5, size=n // 32) * 0.01).astype(np.float16).view(np.uint8).reshape(-1, 2)
    return raw.reshape(-1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=96000)
    ap.add_argument("--cols", type=int, default=4096)
    ap.add_argument("--dtype", default="q4_0")
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--iters", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    a = ap.parse_args()
    typ = G.TYPE_IDS[a.dtype]
    raw = random_blocks(typ, a.rows, a.cols)
    dm = M.DeviceMatrix(raw, typ, a.rows, a.cols)
    copies = max(2, int(300e6 // raw.size) + 1)
    ms = dm.bench(batch=a.batch, n_copies=copies, warmup=a.warmup, iters=a.iters)
    nbytes = raw.size + 4 * a.batch * (a.cols + a.rows)
    peak = 6552.6
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    print(json.dumps({"shape": [a.rows, a.cols], "dtype": a.dtype, "batch": a.batch, "us": ms * 1e3, "GBps": nbytes / ms / 1e6, "frac": nbytes / ms / 1e6 / peak, "copies": copies}))
    dm.close()


if __name__ == "__main__":
    main()
