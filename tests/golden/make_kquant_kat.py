#!/usr/bin/env python3
"""Known-answer vectors for the GGML types the Go engine decodes besides Q4_0 / Q8_0: Q5_0, Q4_K, Q6_K
(go/quant.go:171-276 Q6_K, :282-396 Q4_K, :402-484 Q5_0).

The reference ships no vectors and no producer for these types (its exporters write F32 / F16 / Q8_0 / Q4_0 only), so the
pin is the independent decoder gguf-py 0.19 `gguf.quants.dequantize` (the same decoder that pins Q4_0 / Q8_0 in
make_golden.py):
  * Q5_0: blocks made by gguf-py's own quantizer from a seeded tensor with edge blocks, plus random-byte blocks;
  * Q4_K / Q6_K: gguf-py has no quantizer for them -> random bytes (every 6-bit scale / 4-bit min / high-bit plane
    combination occurs) with the fp16 super-block scale fields kept finite.
Expected values are stored as fp32 bit patterns; oracle and GPU must both reproduce them bit for bit.

Run in the build container:  python tests/golden/make_kquant_kat.py   ->  tests/golden/kquant_kat.npz
"""
import os

import numpy as np
from gguf.constants import GGML_QUANT_SIZES, GGMLQuantizationType as QT
from gguf.quants import dequantize, quantize

HERE = os.path.dirname(os.path.abspath(__file__))


def finite_scales(blocks: np.ndarray, cols) -> None:
    for c in cols:   # high byte of an fp16 field: clear the top exponent bit -> |d| < 2, never inf / NaN
        blocks[:, c] &= 0xBF


def main():
    rng = np.random.default_rng(20261017)
    out = {}
    # ---- Q5_0 (32 elements, 22 bytes: fp16 d | 4 bytes of fifth bits | 16 bytes of nibbles)
    src = rng.standard_normal((96, 32)).astype(np.float32)
    src[0] = 0.0
    src[1] = 0.0; src[1, 5] = 3.25
    src[2] = 0.0; src[2, 17] = -7.5
    src[3] *= 1e-6
    src[4] *= 1e3
    src[5] = np.linspace(-1, 1, 32, dtype=np.float32)
    q = quantize(src, QT.Q5_0).reshape(-1, 22)
    rnd = rng.integers(0, 256, size=(160, 22), dtype=np.uint8)
    finite_scales(rnd, [1])
    raw = np.concatenate([q, rnd]).reshape(-1)
    out["q5_0_bytes"] = raw
    out["q5_0_expect"] = dequantize(raw.reshape(-1, 22), QT.Q5_0).astype(np.float32).reshape(-1).view(np.uint32)
    # ---- Q4_K (256 elements, 144 bytes: fp16 d | fp16 dmin | 12 bytes of packed 6-bit scales / mins | 128 bytes of nibbles)
    bs = GGML_QUANT_SIZES[QT.Q4_K][1]
    raw = rng.integers(0, 256, size=(64, bs), dtype=np.uint8)
    finite_scales(raw, [1, 3])
    out["q4_k_bytes"] = raw.reshape(-1)
    out["q4_k_expect"] = dequantize(raw, QT.Q4_K).astype(np.float32).reshape(-1).view(np.uint32)
    # ---- Q6_K (256 elements, 210 bytes: 128 bytes low nibbles | 64 bytes high 2 bits | 16 int8 scales | fp16 d)
    bs = GGML_QUANT_SIZES[QT.Q6_K][1]
    raw = rng.integers(0, 256, size=(64, bs), dtype=np.uint8)
    finite_scales(raw, [209])
    out["q6_k_bytes"] = raw.reshape(-1)
    out["q6_k_expect"] = dequantize(raw, QT.Q6_K).astype(np.float32).reshape(-1).view(np.uint32)
    for k in ("q5_0", "q4_k", "q6_k"):
        e = out[k + "_expect"].view(np.float32)
        assert np.isfinite(e).all(), k
        print(k, out[k + "_bytes"].size, "bytes ->", e.size, "values; max |v| =", float(np.abs(e).max()))
    np.savez_compressed(os.path.join(HERE, "kquant_kat.npz"), **out)


if __name__ == "__main__":
    main()
