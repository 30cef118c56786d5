"""Decode-mode parity worker (run in a subprocess so that build-time environment switches of the library take effect): one synthetic
model, logits against the oracle at positions that cross the attention split boundaries, and a short greedy stream.

    NL_TILE_POLL=0 python tests/mode_worker.py     # the barrier path of the persistent kernel (what tensor parallelism runs)
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nanollama_b200 import gguf as G
from nanollama_b200 import model as M
from nanollama_b200 import tiers as T
from oracle import oracle as O


def main():
    gf = T.SyntheticGGUF("mini", G.GGML_Q4_0, seed=7, seq_len=224, vocab=2048, layers=6)
    m = M.load_llama_model(gf)
    o = O.OracleModel(gf)
    assert m.decode_path == "decode_tiled_kernel", m.decode_path
    rng = np.random.default_rng(3)
    seq = rng.integers(3, 2048, size=200).astype(np.int32)
    worst = 0.0
    for pos, t in enumerate(seq):
        exp = o.forward(int(t), pos)
        m.forward(int(t), pos)
        if pos in (0, 1, 47, 48, 95, 96, 97, 143, 144, 191, 192, 199):
            rel = float(np.max(np.abs(m.state.logits - exp)) / np.max(np.abs(exp)))
            worst = max(worst, rel)
            assert rel < 2e-5, (pos, rel)
            assert int(np.argmax(m.state.logits)) == int(np.argmax(exp)), pos
    m.reset(); o.reset()
    prompt = np.concatenate([[1], rng.integers(3, 2048, size=15)]).astype(np.int32)
    exp, margins = o.generate_greedy(prompt, 48)
    got = m.generate_greedy(prompt, 48)
    bad = [i for i in range(48) if got[i] != exp[i]]
    assert not bad or margins[bad[0]] < 1e-4, bad[:3]
    m.close(); o.close()
    print(f"MODE_PARITY_OK worst_rel={worst:.2e}")


if __name__ == "__main__":
    main()
