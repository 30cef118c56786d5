#!/usr/bin/env python3
"""Sampling step at the big tier's vocabulary (96k): Engine.generate_tokens with the samplers on the host (logits copied out every
token, the vocabulary sorted by numpy like go/main.go:346-398 sorts it) against nl_sample on the device-resident logits.

    python tools/sample_bench.py [--tier big] [--layers 2] [--tokens 48]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nanollama_b200 import gguf as G
from nanollama_b200 import model as M
from nanollama_b200 import tiers as T
from nanollama_b200.engine import Engine, GenParams


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tier", default="big")
    ap.add_argument("--layers", type=int, default=2)
    ap.add_argument("--tokens", type=int, default=48)
    a = ap.parse_args()
    gf = T.SyntheticGGUF(a.tier, G.GGML_Q4_0, seed=0, seq_len=a.tokens + 32, layers=a.layers)
    m = M.load_llama_model(gf)
    prompt = [1, 17, 4242, 90001 % gf.meta.vocab_size]
    out = {"tier": a.tier, "layers": a.layers, "vocab": gf.meta.vocab_size, "tokens": a.tokens}
    for name, top_p, top_k in (("top_p_0.9", 0.9, 50), ("top_k_50", 1.0, 50)):
        p = GenParams(max_tokens=a.tokens, temperature=0.8, top_p=top_p, top_k=top_k)
        res = {}
        streams = {}
        for dev in (False, True):
            e = Engine(m, eos_id=-1, seed=7, device_sampling=dev)
            e.generate_tokens(prompt, GenParams(max_tokens=4, temperature=0.8, top_p=top_p, top_k=top_k))   # warm
            e = Engine(m, eos_id=-1, seed=7, device_sampling=dev)
            t0 = time.perf_counter()
            streams[dev] = e.generate_tokens(prompt, p)
            res["device" if dev else "host"] = (time.perf_counter() - t0) / a.tokens * 1e3
        res["same_stream"] = streams[False] == streams[True]
        out[name] = {k: (round(v, 3) if isinstance(v, float) else v) for k, v in res.items()}
    # the sampling call alone
    m.forward_device(1, 0)
    for top_p in (0.9, 1.0):
        m.sample(0.8, 50, top_p, 1.0, [], 0.5)
        t0 = time.perf_counter()
        for i in range(50):
            m.sample(0.8, 50, top_p, 1.15, [1, 2, 3], (i + 0.5) / 50)
        out[f"nl_sample_ms_top_p_{top_p}"] = round((time.perf_counter() - t0) / 50 * 1e3, 3)
    print(json.dumps(out))
    m.close()


if __name__ == "__main__":
    main()
