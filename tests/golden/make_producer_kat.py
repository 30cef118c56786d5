#!/usr/bin/env python3
"""Known answers for the fast exact producers: bytes of the REFERENCE's own quantizers on a seeded
131,072-element tensor (4096 blocks, with the edge blocks of make_golden.py mixed in), stored as SHA-256 digests plus the first 512
bytes of each stream.  The source tensor is regenerated from the numpy seed in the test (numpy's PCG64 stream is stable).

    reference functions: scripts/export_gguf.py:85-121 tensor_to_q4_0, :124-159 tensor_to_q8_0,
                         scripts/quantize_gguf.py:183-215 quantize_to_q8_0 (on the fp16-rounded values, as its CLI feeds it)

Run in the build container (needs /root/reference):  python tests/golden/make_producer_kat.py  ->  tests/golden/producer_kat.npz
"""
import hashlib
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REF, "scripts"))
HERE = os.path.dirname(os.path.abspath(__file__))

import export_gguf as EG  # noqa: E402
import quantize_gguf as QG  # noqa: E402

SEED, NBLOCKS = 20261017, 4096


def source(seed=SEED, nblocks=NBLOCKS):
    """Shared with tests/test_gguf_host.py (keep in sync): seeded values with edge blocks."""
    rng = np.random.default_rng(seed)
    t = rng.standard_normal((nblocks, 32)).astype(np.float32)
    t[0] = 0.0
    t[1] = 0.0; t[1, 5] = 3.25
    t[2] = 0.0; t[2, 17] = -7.5
    t[3] *= np.float32(1e-6)
    t[4] *= np.float32(1e3)
    t[5] = np.linspace(-1, 1, 32, dtype=np.float32)
    t[6] = np.float32(6e-8)
    t[7::64] *= np.float32(37.0)
    return t.reshape(-1)


def main():
    x = source()
    tt = torch.from_numpy(x.copy())
    q4 = np.frombuffer(EG.tensor_to_q4_0(tt), dtype=np.uint8)
    q8 = np.frombuffer(EG.tensor_to_q8_0(tt), dtype=np.uint8)
    h = x.astype(np.float16).astype(np.float32)
    q8r = np.frombuffer(QG.quantize_to_q8_0([float(v) for v in h.tolist()]), dtype=np.uint8)
    out = {"seed": np.int64(SEED), "nblocks": np.int64(NBLOCKS)}
    for k, v in (("q4_0", q4), ("q8_0", q8), ("q8_0_requant", q8r)):
        out[k + "_sha256"] = np.frombuffer(hashlib.sha256(v.tobytes()).digest(), dtype=np.uint8)
        out[k + "_head"] = v[:512].copy()
        out[k + "_nbytes"] = np.int64(v.size)
        print(k, v.size, hashlib.sha256(v.tobytes()).hexdigest())
    np.savez_compressed(os.path.join(HERE, "producer_kat.npz"), **out)


if __name__ == "__main__":
    main()
