#!/usr/bin/env python3
"""Small end-to-end run for compute-sanitizer (memcheck / initcheck / synccheck): the reference-written tiny GGUFs through every kernel
family -- the persistent tiled kernel (polled; NL_TILE_POLL=0 for the barrier mode), the per-matrix chain (F16), the one-pass
tensor-core prefill, batched decode, the device sampler, the operator hooks -- with a parity check against the oracle at the end so
that a sanitizer-clean run is also a correct one.

    compute-sanitizer --tool memcheck python tools/sanitizer_target.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nanollama_b200 import gguf as G
from nanollama_b200 import model as M
from oracle import oracle as O


def main():
    gold = np.load(os.path.join(ROOT, "tests", "golden", "golden_logits.npz"))
    toks = gold["tokens"]
    worst = 0.0
    for name in ("tiny_gqa_q4_0", "tiny_gqa_q8_0", "tiny_gqa_f16"):
        gf = G.load_gguf(os.path.join(ROOT, "tests", "golden", name + ".gguf"))
        m = M.load_llama_model(gf, max_batch=2)
        o = O.OracleModel(gf)
        for pos, t in enumerate(toks[:6]):
            m.forward(int(t), pos)
            exp = o.forward(int(t), pos)
            worst = max(worst, float(np.abs(m.state.logits - exp).max() / np.abs(exp).max()))
        got = m.generate_greedy(toks[:4], 6)
        exp_s, _ = o.generate_greedy(toks[:4], 6)
        assert np.array_equal(got, exp_s), (name, got, exp_s)
        m.reset()
        m.prefill(np.resize(toks, 20))                       # one-pass prefill (tcgen05 GEMMs + tensor-core attention)
        m.forward_batch([int(toks[0]), int(toks[1])], [0, 0])
        m.sample(0.8, 5, 0.9, 1.1, [1, 2, 3], 0.37)
        print(f"[sanitizer target] {name}: path {m.decode_path}, logits max-rel {worst:.2e}")
        m.close()
    raw = G.quantize_q4_0(np.random.default_rng(0).standard_normal((64, 128)).astype(np.float32))
    M.dequant(G.GGML_Q4_0, raw, 64 * 128)
    M.matmul_dispatch(raw, G.GGML_Q4_0, np.ones(128, np.float32), 64, 128)
    assert worst < 1e-3, worst
    print("SANITIZER_TARGET_OK")


if __name__ == "__main__":
    main()
