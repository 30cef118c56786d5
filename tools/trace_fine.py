#!/usr/bin/env python3
"""Fine-grained view of the tiled decode kernel from the clock64 sub-stamps (NL_TRACE=<f> also writes <f>.ck): SM cycles spent
between consecutive steps inside one CTA, averaged over CTAs and layers, per phase kind.  Intervals never mix SMs (clock64 is per SM).

    NL_TRACE=trace.bin python bench.py --steps 1 --warmup 1 ...;  python tools/trace_fine.py trace.bin.ck [--mhz 1965]

Slots (nl_tile.cu, TL_CK / CK_AT), GEMV phases:
  0 prologue entry (tid 0)   1 first look issued + block barrier   2 tid 0's first item valid   3 tid 0 converted its items
  4 fragments complete (block barrier passed)   5 tid 0's warp streamed its band   6 warp 15 converted its items
  7 warp 15 streamed its band   8 finishing warp enters the phase   9 last slot consumed by all math warps   10 outputs stored
attention phases (tid 0, the CTA's first item):
  0 entry   1 prefetch issued + barrier   2 q/k/v polled, RoPE written   3 block barrier (+QK-norm)   4 K/V rows in smem + barrier
  5 scores + barrier   6 softmax + barrier   7 PV + barrier   8 PV partials + barrier   9 fold + stores   10 last barrier
"""
import argparse

import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("path")
    ap.add_argument("--mhz", type=float, default=1965.0)
    a = ap.parse_args()
    hdr = np.fromfile(a.path, dtype=np.int32, count=2)
    G, P = int(hdr[0]), int(hdr[1])
    t = np.fromfile(a.path, dtype=np.uint64, offset=8).reshape(G, P, 16).astype(np.float64)
    t[t == 0] = np.nan
    kinds = ["qkv", "attn", "o", "gate/up", "down"]
    name = lambda p: "lm_head" if p == P - 1 else kinds[p % 5]
    us = lambda c: c / a.mhz

    def stat(x):
        x = x[np.isfinite(x)]
        return (float(np.mean(x)), float(np.percentile(x, 95))) if x.size else (float("nan"), float("nan"))

    gemv_iv = [("entry -> first look + barrier", 0, 1), ("-> tid0 first item valid (poll)", 1, 2), ("-> tid0 converted", 2, 3),
               ("-> fragments complete (barrier)", 3, 4), ("   (warp 15 converted -> fragments complete)", 6, 4), ("-> tid0's warp streamed", 4, 5),
               ("   (fragments complete -> warp 15 streamed)", 4, 7), ("tid0 streamed -> finisher sees last slot", 5, 9), ("-> outputs stored", 9, 10)]
    attn_iv = [("entry -> prefetch + barrier", 0, 1), ("-> q/k/v polled + RoPE", 1, 2), ("-> barrier (+QK-norm)", 2, 3), ("-> K/V in smem + barrier", 3, 4),
               ("-> scores + barrier", 4, 5), ("-> softmax + barrier", 5, 6), ("-> PV + barrier", 6, 7), ("-> PV partials + barrier", 7, 8),
               ("-> fold + stores", 8, 9), ("-> last barrier", 9, 10)]
    for k in ["qkv", "attn", "o", "gate/up", "down", "lm_head"]:
        ps = [p for p in range(1, P) if name(p) == k]
        if not ps:
            continue
        print(f"\n## {k} ({len(ps)} phases x {G} CTAs): mean / p95, us at {a.mhz:.0f} MHz (cycles)")
        for label, i, j in (attn_iv if k == "attn" else gemv_iv):
            d = np.concatenate([(t[:, p, j] - t[:, p, i]).ravel() for p in ps])
            m, q = stat(d)
            print(f"  {label:52s} {us(m):7.3f} / {us(q):7.3f}   ({m:8.0f} cyc)")
        # CTA-local gap between the end of the phase before (math warps) and this phase's entry
        d = np.concatenate([(t[:, p, 0] - np.fmax(t[:, p - 1, 5], t[:, p - 1, 10] if name(p - 1) == "attn" else t[:, p - 1, 5])).ravel() for p in ps])
        m, q = stat(d)
        print(f"  {'(previous phase done on this CTA -> entry)':52s} {us(m):7.3f} / {us(q):7.3f}   ({m:8.0f} cyc)")


if __name__ == "__main__":
    main()
