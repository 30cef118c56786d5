"""Prefill throughput of the tcgen05 GEMM path: tok/s and fraction of the tensor roofline (useful FLOPs)."""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nanollama_b200 import gguf as G, model as M, tiers as T

ap = argparse.ArgumentParser()
ap.add_argument("--tier", default="goldie"); ap.add_argument("--dtype", default="q4_0"); ap.add_argument("--tokens", type=int, default=2047)
ap.add_argument("--iters", type=int, default=3)
a = ap.parse_args()
typ = G.TYPE_IDS[a.dtype]
gf = T.SyntheticGGUF(a.tier, typ, seed=0, seq_len=2048)
m = M.load_llama_model(gf)
rng = np.random.default_rng(0)
toks = np.concatenate([[1], rng.integers(3, gf.meta.vocab_size, size=a.tokens - 1)]).astype(np.int32)
m.prefill(toks)  # warm-up (allocates the workspace)
ts = []
for _ in range(a.iters):
    m.reset(); t0 = time.perf_counter(); m.prefill(toks); ts.append(time.perf_counter() - t0)
t = min(ts)
meta = gf.meta
kvd = meta.num_kv_heads * meta.head_dim
layer_params = 2 * meta.embed_dim ** 2 + 2 * kvd * meta.embed_dim + 3 * meta.embed_dim * meta.interm_size
Tn = a.tokens
flops = 2 * Tn * layer_params * meta.num_layers + 2 * meta.vocab_size * meta.embed_dim + meta.num_layers * 4 * meta.embed_dim * Tn * (Tn + 1) / 2
peak = 1628.7
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["bf16_tflops"]
except Exception:
    pass
print(json.dumps({"metric": "prefill tok/s", "tier": a.tier, "dtype": a.dtype, "tokens": Tn, "ms": t * 1e3, "value": Tn / t, "useful_tflops": flops / t / 1e12,
                  "tensor_roofline_frac_useful": flops / t / 1e12 / peak, "issued_tflops_3x_split": 3 * 2 * Tn * layer_params * meta.num_layers / t / 1e12,
                  "peak_tflops": peak, "note": "useful FLOPs per SURVEY 8d; the split-bf16 scheme issues 3 MMAs per useful one"}))
m.close()
