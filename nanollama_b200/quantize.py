"""Fast exact re-quantizer: the vectorised mirror of scripts/quantize_gguf.py, plus the `--dtype q4_0` path the reference lacks.

    python -m nanollama_b200.quantize <input.gguf> <output.gguf> [--dtype q8_0|q4_0]

--dtype q8_0 writes the SAME FILE, byte for byte, as `python scripts/quantize_gguf.py in out` (quantize_gguf.py:262-391): 1-D
tensors become F32, every other F16 / F32 tensor with a multiple of 32 elements becomes Q8_0 with the float64 arithmetic of
quantize_to_q8_0 (:183-215), anything else stays F32; metadata is re-serialised with the reference's array-type inference
(:398-436: non-negative int arrays become uint32); tensor order, dims and the 32-byte alignment are kept.
--dtype q4_0 does the same walk with the Q4_0 block encoder of scripts/export_gguf.py:85-121 (tensor_to_q4_0: scale = amax / 8,
zero block -> 1.0, round-half-even), whose CLI does not accept q4_0 (export_gguf.py:432).  The reference's producers are pure
Python (0.1-0.6 M elements/s); this one is numpy (tens of M elements/s) and is what makes goldie / big files practical.
"""
from __future__ import annotations

import argparse
import struct
import sys

import numpy as np

from . import gguf as G


def _array_elem_type(values) -> int:
    """quantize_gguf.py:_write_array_value: element type inferred from the content."""
    if not values:
        return G.T_UINT32
    first = values[0]
    if isinstance(first, str):
        return G.T_STRING
    if isinstance(first, float):
        return G.T_FLOAT32
    if isinstance(first, bool):
        return G.T_UINT32
    if isinstance(first, int):
        return G.T_INT32 if any(v < 0 for v in values) else G.T_UINT32
    if isinstance(first, list):
        return G.T_ARRAY
    return G.T_UINT32


def _to_f32(raw: np.ndarray, ggml_type: int, n: int) -> np.ndarray:
    if ggml_type == G.GGML_F16:
        return raw.view(np.float16)[:n].astype(np.float32)
    if ggml_type == G.GGML_F32:
        return raw.view(np.float32)[:n]
    raise ValueError(f"Cannot quantize type {ggml_type}")


def requantize(src_path: str, dst_path: str, dtype: str = "q8_0", verbose: bool = True) -> None:
    gf = G.load_gguf(src_path)
    target = {"q8_0": G.GGML_Q8_0, "q4_0": G.GGML_Q4_0}[dtype]
    w = G.GGUFWriter(dst_path)
    w.version = gf.version
    for key, value in gf.meta.kv.items():
        vtype = gf.meta.kv_types[key]
        if vtype == G.T_ARRAY:
            values = list(value)
            w.add_array(key, _array_elem_type(values), values)
        else:
            w.kv.append((key, vtype, value))
    for name, info in gf.tensors_in_file_order():
        raw, _ = gf.get_tensor(name)
        raw = np.ascontiguousarray(raw)
        n = info.n_elements
        shape = tuple(reversed(info.dims))
        if info.ndims == 1:      # norms stay F32 (quantize_gguf.py:289-299)
            out = _to_f32(raw, info.type, n).astype(np.float32).view(np.uint8) if info.type == G.GGML_F16 else raw
            w.add_tensor_raw(name, out.reshape(-1), G.GGML_F32, shape)
            if verbose:
                print(f"  {name:40s} [{n}] -> F32 (norm)")
            continue
        vals = _to_f32(raw, info.type, n)
        if n % 32 != 0:
            w.add_tensor_raw(name, np.ascontiguousarray(vals, dtype=np.float32).view(np.uint8).reshape(-1), G.GGML_F32, shape)
            if verbose:
                print(f"  WARNING: {name} ({n}) not block-compatible, keeping F32")
            continue
        out = G.quantize_q8_0(vals, flavor="requant") if target == G.GGML_Q8_0 else G.quantize_q4_0(vals)
        w.add_tensor_raw(name, out, target, shape)
        if verbose:
            print(f"  {name:40s} [{'x'.join(str(d) for d in shape)}] -> {dtype.upper()} ({raw.size / out.size:.1f}x smaller)")
    w.write()


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description="GGUF F16/F32 -> Q8_0 / Q4_0 (byte-exact mirror of scripts/quantize_gguf.py, plus q4_0)")
    ap.add_argument("input")
    ap.add_argument("output")
    ap.add_argument("--dtype", default="q8_0", choices=["q8_0", "q4_0"])
    ap.add_argument("-q", "--quiet", action="store_true")
    a = ap.parse_args(argv)
    requantize(a.input, a.output, a.dtype, verbose=not a.quiet)
    return 0


if __name__ == "__main__":
    sys.exit(main())
