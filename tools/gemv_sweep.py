#!/usr/bin/env python3
"""BASELINE.json config 5: dequant-GEMV microbench sweep -- every distinct projection shape of the named tiers x {Q4_0, Q8_0, F16}
x batch {1, 4, 16}, L2-cold replicas, GB/s against the measured HBM peak.  Writes a markdown table.

    python tools/gemv_sweep.py [--tiers mini,goldie,big] [--out gpurun_out/gemv_sweep.md]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nanollama_b200 import gguf as G
from nanollama_b200 import model as M
from nanollama_b200 import tiers as T
from tools.gemv_bench import random_blocks


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tiers", default="mini,goldie,big")
    ap.add_argument("--dtypes", default="q4_0,q8_0,f16")
    ap.add_argument("--batches", default="1,4,16")
    ap.add_argument("--out", default="gpurun_out/gemv_sweep.md")
    a = ap.parse_args()
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    rows_out = []
    for tier in a.tiers.split(","):
        m = T.tier_meta(tier)
        kvd = m.num_kv_heads * m.head_dim
        shapes = {"q/o": (m.embed_dim, m.embed_dim), "k/v": (kvd, m.embed_dim), "gate/up": (m.interm_size, m.embed_dim),
                  "down": (m.embed_dim, m.interm_size), "lm_head": (m.vocab_size, m.embed_dim)}
        for name, (r, c) in shapes.items():
            for dt in a.dtypes.split(","):
                typ = G.TYPE_IDS[dt]
                raw = random_blocks(typ, r, c)
                dm = M.DeviceMatrix(raw, typ, r, c)
                copies = max(2, int(300e6 // raw.size) + 1)
                for b in [int(x) for x in a.batches.split(",")]:
                    ms = dm.bench(batch=b, n_copies=copies, warmup=3, iters=20)
                    nbytes = raw.size + 4 * b * (r + c)
                    rows_out.append((tier, name, r, c, dt, b, ms * 1e3, nbytes / ms / 1e6, nbytes / ms / 1e6 / peak))
                dm.close()
    with open(a.out, "w") as f:
        f.write(f"# Dequant-GEMV sweep (config 5): L2-cold replicas, CUDA events, measured HBM peak {peak:.1f} GB/s\n\n")
        f.write("Launch-to-launch time of a stand-alone GEMV (each launch pays ~8 us of launch + prologue + tail: small matrices are latency-bound;\n"
                "inside the persistent decode kernel the same code streams at the roofline, see r01_tiled_phase_trace.md).\n\n")
        f.write("| tier | matrix | rows | cols | type | batch | us | GB/s | frac of HBM peak |\n|---|---|---|---|---|---|---|---|---|\n")
        for t in rows_out:
            f.write(f"| {t[0]} | {t[1]} | {t[2]} | {t[3]} | {t[4]} | {t[5]} | {t[6]:.1f} | {t[7]:.0f} | {t[8]:.3f} |\n")
    print(open(a.out).read())


if __name__ == "__main__":
    main()
