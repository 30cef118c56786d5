#!/usr/bin/env python3
"""A/B of decode-kernel variants selected by environment variables, one subprocess per variant on the same synthetic model.

    python tools/decode_ab.py --tier big --layers 10 --variants "NL_TILE_POLL=0;NL_TILE_POLL=1;NL_TILE_POLL=1,NL_ATT_CHUNK=32"

Every variant decodes the same greedy stream (token ids and last logits are compared with the FIRST variant, which should be the
one whose parity against the oracle is already established) and is timed with nl_bench_decode.  --trace DIR additionally writes an
NL_TRACE dump per variant (tools/trace_summary.py reads it).
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(a):
    import numpy as np
    from nanollama_b200 import gguf as G
    from nanollama_b200 import model as M
    from nanollama_b200 import tiers as T
    typ = G.TYPE_IDS[a.dtype]
    gf = T.SyntheticGGUF(a.tier, typ, seed=0, seq_len=a.pos + a.steps + 80, layers=a.layers or None)
    m = M.load_llama_model(gf)
    rng = np.random.default_rng(5)
    prompt = np.concatenate([[1], rng.integers(3, gf.meta.vocab_size, size=15)]).astype(np.int32)
    toks = m.generate_greedy(prompt, a.gen)
    m.forward(int(toks[-1]), len(prompt) + len(toks) - 1)   # one more step through nl_forward: the logits the variants are compared on
    logits = np.array(m.state.logits, dtype=np.float32).reshape(-1)[: gf.meta.vocab_size].copy()
    m.reset()
    m.bench_decode(1, a.pos, 16)
    ms = min(m.bench_decode(1, a.pos, a.steps) for _ in range(a.reps))
    mid = a.pos + a.steps // 2
    b = T.decode_bytes_per_token(gf.meta, typ, mid)
    np.save(a.out + ".logits.npy", logits)
    json.dump({"tok_s": a.steps / ms * 1e3, "us_tok": ms / a.steps * 1e3, "GBps": b / (ms / a.steps) / 1e6, "tokens": [int(t) for t in toks],
               "path": m.decode_path}, open(a.out + ".json", "w"))
    m.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tier", default="big")
    ap.add_argument("--dtype", default="q4_0")
    ap.add_argument("--layers", type=int, default=0, help="truncate the tier to this many layers (0 = full depth): shorter loads")
    ap.add_argument("--pos", type=int, default=16)
    ap.add_argument("--steps", type=int, default=256)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--gen", type=int, default=48)
    ap.add_argument("--variants", default="NL_TILE_POLL=0;NL_TILE_POLL=1")
    ap.add_argument("--trace", default="")
    ap.add_argument("--timeout", type=int, default=180)
    ap.add_argument("--child", action="store_true")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    if a.child:
        return child(a)
    import numpy as np
    base = None
    rows = []
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    for i, v in enumerate(a.variants.split(";")):
        env = dict(os.environ)
        for kv in filter(None, v.split(",")):
            k, _, val = kv.partition("=")
            env[k.strip()] = val.strip()
        out = os.path.join(ROOT, "gpurun_out", f"ab_{i}")
        if a.trace:
            os.makedirs(a.trace, exist_ok=True)
            env["NL_TRACE"] = os.path.join(a.trace, f"trace_{i}.bin")
        cmd = [sys.executable, os.path.abspath(__file__), "--child", "--out", out, "--tier", a.tier, "--dtype", a.dtype, "--layers", str(a.layers),
               "--pos", str(a.pos), "--steps", str(a.steps), "--reps", str(a.reps), "--gen", str(a.gen)]
        for f in (out + ".json", out + ".logits.npy"):
            if os.path.exists(f):
                os.remove(f)
        try:
            r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=a.timeout)
            err = r.stderr.strip().splitlines()[-1] if r.returncode else ""
        except subprocess.TimeoutExpired:
            err = f"TIMEOUT after {a.timeout}s (hang?)"
        if err or not os.path.exists(out + ".json"):
            rows.append({"variant": v, "error": err or "no output"})
            print(json.dumps(rows[-1]), flush=True)
            continue
        d = json.load(open(out + ".json"))
        lg = np.load(out + ".logits.npy")
        if base is None:
            base = (d["tokens"], lg)
        same = d["tokens"] == base[0]
        rel = float(np.max(np.abs(lg - base[1])) / max(float(np.max(np.abs(base[1]))), 1e-30))
        rows.append({"variant": v, "tok_s": round(d["tok_s"], 1), "us_tok": round(d["us_tok"], 1), "GBps": round(d["GBps"], 1), "path": d["path"],
                     "stream_same_as_first": same, "logits_maxrel_vs_first": rel})
        print(json.dumps(rows[-1]), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
