#!/usr/bin/env python3
"""Batch-1 decode speed as a function of the context position (the KV-cache read grows with pos; SURVEY 8d bytes_tok(pos)).

    python tools/decode_ctx_bench.py --tier big --dtype q4_0 --positions 16,256,1024,1900 --steps 64
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nanollama_b200 import gguf as G
from nanollama_b200 import model as M
from nanollama_b200 import tiers as T


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tier", default="big")
    ap.add_argument("--dtype", default="q4_0")
    ap.add_argument("--positions", default="16,256,1024,1900")
    ap.add_argument("--steps", type=int, default=64)
    a = ap.parse_args()
    typ = G.TYPE_IDS[a.dtype]
    gf = T.SyntheticGGUF(a.tier, typ, seed=0, seq_len=2048)
    m = M.load_llama_model(gf)
    m.reset()
    out = []
    for p in [int(x) for x in a.positions.split(",")]:
        m.bench_decode(1, p, 8)
        ms = m.bench_decode(1, p, a.steps)
        mid = p + a.steps // 2
        b = T.decode_bytes_per_token(gf.meta, typ, mid)
        out.append({"pos": p, "tok_s": a.steps / ms * 1e3, "GBps": b / (ms / a.steps) / 1e6, "bytes_tok": b})
    print(json.dumps({"tier": a.tier, "dtype": a.dtype, "path": m.decode_path, "results": out}))
    m.close()


if __name__ == "__main__":
    main()
