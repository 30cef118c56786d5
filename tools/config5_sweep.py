#!/usr/bin/env python3
"""BASELINE.json config 5, the whole table: every distinct projection shape of every tier x {Q4_0, Q8_0, F16} x decode batch
B in {1, 2, 4, 8, 16, 32, 64} (GB/s of algorithmic bytes against the measured HBM peak) and prefill row counts M in {512, 2048}
(useful TFLOP/s against the measured bf16 tensor peak), L2-cold replicas, CUDA events inside nl_matrix_bench.

    python tools/config5_sweep.py [--tiers nano,micro,...] [--out gpurun_out/config5_sweep.md]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nanollama_b200 import gguf as G
from nanollama_b200 import model as M
from nanollama_b200 import tiers as T
from tools.gemv_bench import random_blocks


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tiers", default="nano,micro,mini,small,goldie,medium,large,big")
    ap.add_argument("--dtypes", default="q4_0,q8_0,f16")
    ap.add_argument("--batches", default="1,2,4,8,16,32,64")
    ap.add_argument("--rows", default="512,2048")
    ap.add_argument("--out", default="gpurun_out/config5_sweep.md")
    a = ap.parse_args()
    pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    hbm, tens = float(pk["hbm_gbs"]), float(pk["bf16_tflops"])
    seen, out = set(), []
    for tier in a.tiers.split(","):
        m = T.tier_meta(tier)
        kvd = m.num_kv_heads * m.head_dim
        shapes = {"q/o": (m.embed_dim, m.embed_dim), "k/v": (kvd, m.embed_dim), "gate/up": (m.interm_size, m.embed_dim),
                  "down": (m.embed_dim, m.interm_size), "lm_head": (m.vocab_size, m.embed_dim)}
        for name, (r, c) in shapes.items():
            for dt in a.dtypes.split(","):
                if (r, c, dt) in seen:
                    continue
                seen.add((r, c, dt))
                typ = G.TYPE_IDS[dt]
                raw = random_blocks(typ, r, c)
                try:
                    dm = M.DeviceMatrix(raw, typ, r, c)
                except Exception as e:
                    out.append((tier, name, r, c, dt, "-", None, None, None, str(e)[:60]))
                    continue
                copies = max(2, int(300e6 // raw.size) + 1)
                for b in [int(x) for x in a.batches.split(",")]:
                    try:
                        ms = dm.bench(batch=b, n_copies=copies, warmup=2, iters=10)
                        nbytes = raw.size + 4 * b * (r + c)
                        out.append((tier, name, r, c, dt, f"B={b}", ms * 1e3, f"{nbytes / ms / 1e6:.0f} GB/s", nbytes / ms / 1e6 / hbm, "hbm"))
                    except Exception as e:
                        out.append((tier, name, r, c, dt, f"B={b}", None, None, None, str(e)[:60]))
                if c % 64 == 0:
                    for mrows in [int(x) for x in a.rows.split(",")]:
                        try:
                            ms = dm.bench(batch=mrows, n_copies=min(copies, 4), warmup=1, iters=5)
                            tf = 2.0 * mrows * r * c / ms / 1e9
                            out.append((tier, name, r, c, dt, f"M={mrows}", ms * 1e3, f"{tf:.1f} TFLOP/s", tf / tens, "tensor"))
                        except Exception as e:
                            out.append((tier, name, r, c, dt, f"M={mrows}", None, None, None, str(e)[:60]))
                dm.close()
    with open(a.out, "w") as f:
        f.write(f"# Config 5 sweep: dequant-GEMV (B = decode batch) against the measured HBM peak {hbm:.1f} GB/s, dequant-GEMM (M = prompt rows) against the "
                f"measured bf16 tensor peak {tens:.1f} TFLOP/s\n\n")
        f.write("L2-cold replicas, CUDA events around back-to-back launches (launch-to-launch: every launch pays its own prologue and tail, so small "
                "matrices are latency-bound; inside the persistent decode kernel the same tiles stream without those).  B = 1 on Q4_0 / Q8_0 runs the "
                "tiled tensor-core GEMV of the decode kernel as a one-phase launch; 2 <= B <= 15 the CUDA-core batch GEMV; B >= 16 and M the tcgen05 GEMM.\n\n")
        f.write("| tier | matrix | rows | cols | type | case | us | achieved | fraction of peak | bound |\n|---|---|---|---|---|---|---|---|---|---|\n")
        for t in out:
            if t[6] is None:
                f.write(f"| {t[0]} | {t[1]} | {t[2]} | {t[3]} | {t[4]} | {t[5]} | - | - | - | {t[9]} |\n")
            else:
                f.write(f"| {t[0]} | {t[1]} | {t[2]} | {t[3]} | {t[4]} | {t[5]} | {t[6]:.1f} | {t[7]} | {t[8]:.3f} | {t[9]} |\n")
    print(f"{len(out)} rows -> {a.out}")


if __name__ == "__main__":
    main()
