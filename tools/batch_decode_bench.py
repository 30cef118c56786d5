#!/usr/bin/env python3
"""Small-batch decode (BASELINE config 3: goldie Q4_0, B = 1..64 independent sequences, same weights): tokens per second through
nl_forward_batch with the logits left on the device, and the bytes-per-step roofline fraction (the weights are read once per step).

    python tools/batch_decode_bench.py [--tier goldie] [--batches 1,2,4,8,16,32,64] [--steps 24]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nanollama_b200 import capi
from nanollama_b200 import gguf as G
from nanollama_b200 import model as M
from nanollama_b200 import tiers as T


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tier", default="goldie")
    ap.add_argument("--dtype", default="q4_0")
    ap.add_argument("--batches", default="1,2,4,8,16,32,64")
    ap.add_argument("--steps", type=int, default=24)
    a = ap.parse_args()
    typ = G.TYPE_IDS[a.dtype]
    bs = [int(x) for x in a.batches.split(",")]
    gf = T.SyntheticGGUF(a.tier, typ, seed=0, seq_len=a.steps + 24)
    m = M.load_llama_model(gf, max_batch=max(bs))
    peak = 6552.6
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    wbytes = T.decode_bytes_per_token(gf.meta, typ, 0)   # weights + one position of KV: the per-step floor at small context
    out = []
    L = capi.lib()
    for B in bs:
        toks = np.arange(3, 3 + B, dtype=np.int32)
        for warm in (True, False):
            m.reset()
            n = 4 if warm else a.steps
            t0 = time.perf_counter()
            for s in range(n):
                pos = np.full(B, s, dtype=np.int32)
                capi.check(L.nl_forward_batch(m._h, B, capi.ptr(toks), capi.ptr(pos), None))
            dt = (time.perf_counter() - t0) / n
        out.append({"B": B, "ms_step": round(dt * 1e3, 3), "tok_s": round(B / dt, 1), "frac_of_weight_roofline": round(wbytes / dt / 1e9 / peak, 4)})
    print(json.dumps({"tier": a.tier, "dtype": a.dtype, "steps": a.steps, "path_b1": m.decode_path, "results": out}))
    m.close()


if __name__ == "__main__":
    main()
