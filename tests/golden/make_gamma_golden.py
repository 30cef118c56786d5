"""Writes tests/golden/gamma_ref.npz with the reference's own writer (scripts/extract_gamma.py:141-200 save_gamma_npz) and
tests/golden/gamma_ref_dense.npy, the dense diff it was made from.  Run in the build container only:

    PYTHONPATH=/root/reference python tests/golden/make_gamma_golden.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference")
sys.path.insert(0, "/root/reference/scripts")
from extract_gamma import save_gamma_npz  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    g = torch.Generator().manual_seed(7)
    vocab, dim = 256, 128                      # the tiny golden models' vocabulary and width
    diff = torch.zeros(vocab, dim)
    for tok in (3, 17, 18, 99, 200, 255):
        cols = torch.randperm(dim, generator=g)[:40]
        diff[tok, cols] = torch.randn(40, generator=g) * 0.05
    diff[17, 5] = 1e-9                          # below the writer's sparsity threshold: dropped
    gamma = {"tok_embeddings.weight": {"diff": diff, "norm": float(diff.norm())},
             "layers.0.attention.wq.weight": {"diff": torch.randn(8, 8, generator=g) * 0.01, "norm": 0.1}}   # a second key, like a real file
    out = os.path.join(HERE, "gamma_ref.npz")
    save_gamma_npz(gamma, out)
    np.save(os.path.join(HERE, "gamma_ref_dense.npy"), diff.numpy())
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
