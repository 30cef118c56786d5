#!/usr/bin/env python3
"""Numerics of the tiled kernel's activation fragments, on the CPU: the input x enters the tensor cores as fp16 hi + lo terms scaled by a
power of two per quant block (today: one scale per 32-element block; round-2 plan: one per 16-element half block, the unit a
producer's finishing step holds -- DESIGN section 6).  Emulates the split, exact fp16 x nibble products and fp32 accumulation, and
reports the error of W.x against a float64 reference for both granularities, next to the error of the plain fp32 sequential dot the
reference engine computes (go/quant.go:45-94).

    python tools/frag_numerics.py [--rows 256] [--cols 4096] [--trials 4]
"""
import argparse
import json

import numpy as np


def split_hi_lo(y, group):
    """y [cols] fp32 -> (hi, lo) fp16 values of y * S_g and the scales S_g (power of two per `group` elements, max |y| * S in [2^10, 2^11))."""
    yg = y.reshape(-1, group)
    mx = np.abs(yg).max(axis=1)
    e = np.where(mx > 0, np.floor(np.log2(np.maximum(mx, 1e-38))), 0.0)
    S = np.exp2(10.0 - e).astype(np.float32)            # max|y| * S in [2^10, 2^11)
    v = (yg * S[:, None]).astype(np.float32)
    hi = v.astype(np.float16)
    lo = (v - hi.astype(np.float32)).astype(np.float16)
    return hi.astype(np.float64), lo.astype(np.float64), S.astype(np.float64)


def gemv_frag(q, d, y, group):
    """q [rows, cols] ints in [-8, 7], d [rows, cols/32] block scales: sum over blocks of d * (sum_j q_j * (hi_j + lo_j)) / S, the per-group
    partial dots accumulated in fp32 like the MMA accumulators (products exact)."""
    rows, cols = q.shape
    hi, lo, S = split_hi_lo(y, group)
    part = np.einsum("rgj,gj->rg", q.reshape(rows, -1, group).astype(np.float64), hi + lo)     # exact in float64 (small ints x fp16)
    part = part.astype(np.float32).astype(np.float64) / S[None, :]                                  # fp32 accumulator, then the 2^k / S factor
    per_block = part.reshape(rows, cols // 32, 32 // group).sum(axis=2).astype(np.float32)
    out = np.zeros(rows, np.float32)
    for b in range(cols // 32):                                                                    # fp32 FMA chain over the blocks
        out = (out + d[:, b].astype(np.float32) * per_block[:, b]).astype(np.float32)
    return out


def gemv_ref_f32(q, d, x):
    """The reference's arithmetic: per block an fp32 sequential dot, then * d, accumulated in fp32 (go/quant.go:73-94)."""
    rows, cols = q.shape
    out = np.zeros(rows, np.float32)
    for b in range(cols // 32):
        s = np.zeros(rows, np.float32)
        for j in range(32):
            s = (s + q[:, 32 * b + j].astype(np.float32) * np.float32(x[32 * b + j])).astype(np.float32)
        out = (out + s * d[:, b].astype(np.float32)).astype(np.float32)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=256)
    ap.add_argument("--cols", type=int, default=4096)
    ap.add_argument("--trials", type=int, default=4)
    a = ap.parse_args()
    rng = np.random.default_rng(0)
    res = {"rows": a.rows, "cols": a.cols, "cases": []}
    for t in range(a.trials):
        q = rng.integers(-8, 8, size=(a.rows, a.cols))
        d = (rng.uniform(0.5, 1.5, size=(a.rows, a.cols // 32)) * 0.01).astype(np.float16).astype(np.float64)
        x = rng.standard_normal(a.cols).astype(np.float32)
        if t % 2:                                       # heavy-tailed activations: a few outliers per block group
            x[rng.integers(0, a.cols, size=a.cols // 64)] *= 300.0
        exact = (q.astype(np.float64).reshape(a.rows, -1, 32) * x.astype(np.float64).reshape(-1, 32)[None]).sum(axis=2)
        exact = (exact * d).sum(axis=1)
        scale = np.abs(exact).max()
        case = {"outliers": bool(t % 2)}
        for name, y in (("frag32", gemv_frag(q, d, x, 32)), ("frag16", gemv_frag(q, d, x, 16)), ("ref_fp32", gemv_ref_f32(q, d, x))):
            case[name] = float(np.abs(y.astype(np.float64) - exact).max() / scale)
        res["cases"].append(case)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
