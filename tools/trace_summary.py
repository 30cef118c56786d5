#!/usr/bin/env python3
"""Summarise an NL_TRACE dump of the persistent decode kernel: where one token's time goes, per phase kind.

    NL_TRACE=trace.bin python bench.py --steps 1 --warmup 1 ...;  python tools/trace_summary.py trace.bin

File: int32 grid, int32 n_phases, then uint64 [grid][n_phases][8] globaltimer stamps (ns) of the last token:
  0 barrier wait begins   1 barrier passed   2 input fragments ready   3 last slot consumed   4 phase outputs published (arrive)
"""
import sys

import numpy as np


def main(path):
    hdr = np.fromfile(path, dtype=np.int32, count=2)
    G, P = int(hdr[0]), int(hdr[1])
    t = np.fromfile(path, dtype=np.uint64, offset=8).reshape(G, P, 8).astype(np.float64)
    t[t == 0] = np.nan
    kinds = ["qkv", "attn", "o", "gate/up", "down"]
    name = lambda p: "lm_head" if p == P - 1 else kinds[p % 5]
    end = np.nanmax(t[:, :, 4], axis=0)                       # phase complete (last CTA published)
    start = np.concatenate([[np.nanmin(t[:, 0, 2]) if np.isfinite(np.nanmin(t[:, 0, 2])) else end[0]], end[:-1]])
    rows = {}
    for p in range(1, P):
        d = rows.setdefault(name(p), [])
        w0, w1, f, c, a = t[:, p, 0], t[:, p, 1], t[:, p, 2], t[:, p, 3], t[:, p, 4]
        d.append([end[p] - end[p - 1],                          # phase duration on the critical path
                  np.nanmean(w1 - end[p - 1]),                  # last arrive of p-1 -> barrier seen (mean over CTAs)
                  np.nanmean(f - w1), np.nanmax(f - w1),        # prologue
                  np.nanmean(c - f), np.nanmax(c - f),          # consume
                  np.nanmean(a - c), np.nanmax(a - c),          # finish + publish
                  np.nanmean(w1 - w0),                          # time spent waiting at the barrier (idle)
                  np.nanmean(t[:, p, 5] - w1), np.nanmean(t[:, p, 6] - t[:, p, 5]), np.nanmean(t[:, p, 7] - t[:, p, 6]),   # attention: inputs, own positions, combine
                  np.nanmean(t[:, p, 5] - c), np.nanmean(a - t[:, p, 5]), np.nanmean(t[:, p, 6] - a),   # GEMV finisher: all warps done, sums+stores, fence+arrive
                  np.nanmean(f - t[:, p, 5])])                    # attention: first pass (stamp 2 - stamp 5)
    print(f"grid {G}, phases {P}; token span {(end[-1] - np.nanmin(t[:, 0, 2])) / 1e3:.1f} us (from first fragments of phase 0)")
    print("| phase | n | duration us | barrier seen after last arrive | prologue mean/max | consume mean/max | finish mean/max | idle at barrier mean |")
    print("|---|---|---|---|---|---|---|---|")
    tot = 0.0
    for k in ["qkv", "attn", "o", "gate/up", "down", "lm_head"]:
        if k not in rows:
            continue
        a = np.array(rows[k]) / 1e3
        m = np.nanmean(a, axis=0)
        tot += np.nansum(a[:, 0])
        print(f"| {k} | {len(a)} | {m[0]:.2f} | {m[1]:.2f} | {m[2]:.2f} / {m[3]:.2f} | {m[4]:.2f} / {m[5]:.2f} | {m[6]:.2f} / {m[7]:.2f} | {m[8]:.2f} |", end="")
        if k == "attn":
            print(f"  [attention: q/k/v in smem {m[9]:.2f}, own positions {m[10]:.2f} (first pass {m[15]:.2f}), fold+store {m[11]:.2f}]")
        else:
            print(f"  [finisher: last warp after tid0 {m[12]:.2f}, sums+stores {m[13]:.2f}, fence+arrive {m[14]:.2f}]")
    print(f"sum of phase durations: {tot:.1f} us")


if __name__ == "__main__":
    main(sys.argv[1])
