"""ctypes binding of libnanollama_cuda.so (include/nanollama_cuda.h).  No torch types cross this boundary.

This is the Python twin of the cgo binding in go/model_cuda.go: same calls, same order.  The library is built in-tree by
``nanollama_b200.build``; loading fails loudly if it is missing — there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
# NL_LIB: a differently built copy of the same library (tools/build_variants.py), for A/B runs of build-time variants
LIB_PATH = os.environ.get("NL_LIB") or os.path.join(PKG, "libnanollama_cuda.so")

NL_OK, NL_ERR_INVALID, NL_ERR_CUDA, NL_ERR_UNSUPPORTED, NL_ERR_STATE, NL_ERR_OOM = 0, -1, -2, -3, -4, -5

SLOTS = {"token_embd": 0, "output_norm": 1, "output": 2, "attn_norm": 3, "ffn_norm": 4, "attn_q": 5, "attn_k": 6, "attn_v": 7,
         "attn_output": 8, "ffn_gate": 9, "ffn_up": 10, "ffn_down": 11, "attn_q.bias": 12, "attn_k.bias": 13, "attn_v.bias": 14,
         "attn_output.bias": 15}


class NlConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("n_layers", "embed_dim", "n_heads", "n_kv_heads", "head_dim", "vocab_size", "seq_len", "interm_size")] + \
               [("rms_norm_eps", C.c_float), ("rope_theta", C.c_float)] + \
               [(n, C.c_int32) for n in ("qk_norm", "rope_conjugate", "device", "tp_rank", "tp_size", "max_batch")]


class NlError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libnanollama_cuda error {code}: {msg}")
        self.code = code
        self.msg = msg


# every symbol include/nanollama_cuda.h declares: (restype, argtypes)
_vp, _i32, _i64, _u32, _sz = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_size_t
SIGNATURES = {
    "nl_last_error": (C.c_char_p, []),
    "nl_abi_version": (C.c_int, []),
    "nl_device_count": (C.c_int, []),
    "nl_create": (C.c_int, [C.POINTER(NlConfig), C.POINTER(_vp)]),
    "nl_upload_tensor": (C.c_int, [_vp, C.c_int, C.c_int, _u32, _i64, _i64, _vp, _sz]),
    "nl_set_gamma": (C.c_int, [_vp, _vp, _i32, _vp]),
    "nl_finalize": (C.c_int, [_vp]),
    "nl_destroy": (None, [_vp]),
    "nl_get_config": (C.c_int, [_vp, C.POINTER(NlConfig)]),
    "nl_forward": (C.c_int, [_vp, _i32, _i32, _vp]),
    "nl_reset": (C.c_int, [_vp]),
    "nl_get_logits": (C.c_int, [_vp, _vp]),
    "nl_generate_greedy": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _vp, C.POINTER(_i32)]),
    "nl_sample": (C.c_int, [_vp, C.c_float, _i32, C.c_float, C.c_float, _vp, _i32, C.c_float, C.POINTER(_i32)]),
    "nl_forward_batch": (C.c_int, [_vp, _i32, _vp, _vp, _vp]),
    "nl_prefill": (C.c_int, [_vp, _vp, _i32, _i32, _vp]),
    "nl_dequant": (C.c_int, [_u32, _vp, _i64, _vp]),
    "nl_matmul": (C.c_int, [_u32, _vp, _i64, _i64, _vp, _i32, _vp]),
    "nl_matrix_create": (C.c_int, [_u32, _vp, _i64, _i64, _i32, C.POINTER(_vp)]),
    "nl_matrix_matmul": (C.c_int, [_vp, _vp, _i32, _vp]),
    "nl_matrix_bench": (C.c_int, [_vp, _i32, _i32, _i32, _i32, C.POINTER(C.c_float)]),
    "nl_matrix_destroy": (None, [_vp]),
    "nl_bench_decode": (C.c_int, [_vp, _i32, _i32, _i32, C.POINTER(C.c_float)]),
    "nl_bench_prefill": (C.c_int, [_vp, _vp, _i32, _i32, C.POINTER(C.c_float), C.POINTER(_i32)]),
    "nl_launches_per_token": (C.c_int, [_vp]),
    "nl_weight_bytes": (_i64, [_vp]),
    "nl_decode_path": (C.c_char_p, [_vp]),
    "nl_tp_export_handle": (C.c_int, [_vp, _vp]),
    "nl_tp_import_handles": (C.c_int, [_vp, _vp, _i32]),
}

_lib = None


def lib():
    """Load the shared library (once).  Raises if it has not been built — never falls back to anything."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -m nanollama_b200.build` (there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the .so does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int):
    if rc != NL_OK:
        raise NlError(rc, lib().nl_last_error().decode("utf-8", "replace"))


def ptr(a) -> C.c_void_p:
    return a.ctypes.data_as(C.c_void_p)
