"""ctypes front of oracle/libnl_oracle.so (the C restatement of the Go engine).

TEST INFRASTRUCTURE ONLY — importable from tests/, __graft_entry__.smoke() and bench.py's CPU legs.
Never import this from nanollama_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_DIR, "libnl_oracle.so")

SLOTS = {"token_embd": 0, "output_norm": 1, "output": 2, "attn_norm": 3, "ffn_norm": 4, "attn_q": 5, "attn_k": 6, "attn_v": 7,
         "attn_output": 8, "ffn_gate": 9, "ffn_up": 10, "ffn_down": 11, "attn_q.bias": 12, "attn_k.bias": 13, "attn_v.bias": 14,
         "attn_output.bias": 15}


class NloConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("n_layers", "embed_dim", "n_heads", "n_kv_heads", "head_dim", "vocab_size", "seq_len", "interm_size")] + \
               [("rms_norm_eps", C.c_float), ("rope_theta", C.c_float), ("qk_norm", C.c_int32), ("rope_conjugate", C.c_int32)]


def build(force: bool = False) -> str:
    src = os.path.join(_DIR, "nl_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _DIR, "-B" if force else "-s"], check=True, stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.nlo_half2float.restype = C.c_float
        L.nlo_half2float.argtypes = [C.c_uint16]
        L.nlo_tensor_bytes.restype = C.c_int64
        L.nlo_tensor_bytes.argtypes = [C.c_int, C.c_int64]
        L.nlo_dequant.argtypes = [C.c_int, C.c_void_p, C.c_int64, C.c_void_p]
        L.nlo_matmul.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int]
        L.nlo_rmsnorm.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float]
        L.nlo_rmsnorm_bare.argtypes = [C.c_void_p, C.c_int, C.c_float]
        L.nlo_softmax.argtypes = [C.c_void_p, C.c_int]
        L.nlo_silu.restype = C.c_float
        L.nlo_silu.argtypes = [C.c_float]
        L.nlo_argmax.argtypes = [C.c_void_p, C.c_int]
        L.nlo_rep_penalty.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_float]
        L.nlo_sample_top_k.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_float]
        L.nlo_sample_top_p.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float]
        L.nlo_model_new.restype = C.c_void_p
        L.nlo_model_new.argtypes = [C.POINTER(NloConfig)]
        L.nlo_model_free.argtypes = [C.c_void_p]
        L.nlo_model_set_tensor.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64]
        L.nlo_model_set_gamma.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.nlo_forward.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.nlo_reset.argtypes = [C.c_void_p]
        for f in ("nlo_logits", "nlo_state_x", "nlo_key_cache", "nlo_value_cache", "nlo_rope_cos", "nlo_rope_sin"):
            getattr(L, f).restype = C.POINTER(C.c_float)
            getattr(L, f).argtypes = [C.c_void_p]
        L.nlo_generate_greedy.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.nlo_init()
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def set_workers(n: int):
    lib().nlo_set_workers(int(n))


def get_workers() -> int:
    return lib().nlo_get_workers()


def dequant(ggml_type: int, raw: np.ndarray, n: int) -> np.ndarray:
    raw = np.ascontiguousarray(raw, dtype=np.uint8)
    out = np.empty(n, dtype=np.float32)
    if lib().nlo_dequant(int(ggml_type), _p(raw), n, _p(out)) != 0:
        raise ValueError(f"unsupported tensor type {ggml_type}")
    return out


def matmul(raw: np.ndarray, ggml_type: int, x: np.ndarray, rows: int, cols: int) -> np.ndarray:
    """matmulDispatch (go/model.go:361)."""
    raw = np.ascontiguousarray(raw, dtype=np.uint8)
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.zeros(rows, dtype=np.float32)
    if lib().nlo_matmul(_p(out), _p(raw), int(ggml_type), _p(x), rows, cols) != 0:
        raise ValueError(f"unsupported matmul type {ggml_type}")
    return out


def rmsnorm(x, w, eps):
    x = np.ascontiguousarray(x, dtype=np.float32)
    w = np.ascontiguousarray(w, dtype=np.float32)
    out = np.empty_like(x)
    lib().nlo_rmsnorm(_p(out), _p(x), _p(w), x.size, eps)
    return out


def softmax(x):
    x = np.array(x, dtype=np.float32, copy=True)
    lib().nlo_softmax(_p(x), x.size)
    return x


def silu(x: float) -> float:
    return lib().nlo_silu(float(x))


def rep_penalty(logits: np.ndarray, recent, penalty: float) -> np.ndarray:
    """go/main.go:177-187 on a copy of `logits`."""
    out = np.ascontiguousarray(logits, dtype=np.float32).copy()
    r = np.ascontiguousarray(recent, dtype=np.int32)
    lib().nlo_rep_penalty(_p(out), out.size, _p(r) if r.size else None, r.size, float(penalty))
    return out


def sample_top_k(logits: np.ndarray, temp: float, top_k: int, r01: float) -> int:
    """sampleTopK, go/main.go:294-343; r01 stands for rng.Float32()."""
    lg = np.ascontiguousarray(logits, dtype=np.float32)
    return int(lib().nlo_sample_top_k(_p(lg), lg.size, float(temp), int(top_k), float(r01)))


def sample_top_p(logits: np.ndarray, temp: float, top_p: float, r01: float) -> int:
    """sampleTopP, go/main.go:346-398; r01 stands for rng.Float32()."""
    lg = np.ascontiguousarray(logits, dtype=np.float32)
    return int(lib().nlo_sample_top_p(_p(lg), lg.size, float(temp), float(top_p), float(r01)))


class OracleModel:
    """LoadLlamaModel + Forward + Reset of the Go engine, on the CPU (go/model.go:121-631)."""

    def __init__(self, gf, rope_conjugate=None, qk_norm=None, seq_len=None):
        m = gf.meta
        self.gf = gf
        c = NloConfig(m.num_layers, m.embed_dim, m.num_heads, m.num_kv_heads, m.head_dim, m.vocab_size,
                      seq_len if seq_len is not None else m.seq_len, m.interm_size, m.rms_norm_eps, m.rope_theta,
                      int(m.qk_norm if qk_norm is None else qk_norm), int(m.rope_conjugate if rope_conjugate is None else rope_conjugate))
        self.L = lib()
        self.h = self.L.nlo_model_new(C.byref(c))
        self.cfg = c
        self.seq_len = min(c.seq_len, 2048)
        self.vocab = m.vocab_size
        self.dim = m.embed_dim
        self._keep = []
        self._gamma = None

        def put(slot, layer, name, optional=False):
            try:
                raw, info = gf.get_tensor(name)
            except KeyError:
                if optional:
                    return False
                raise
            raw = np.ascontiguousarray(raw)
            self._keep.append(raw)
            if self.L.nlo_model_set_tensor(self.h, SLOTS[slot], layer, info.type, _p(raw), info.n_elements) != 0:
                raise ValueError(f"unsupported tensor type {info.type} for {name}")
            return True

        put("token_embd", -1, "token_embd.weight")
        put("output_norm", -1, "output_norm.weight")
        if not put("output", -1, "output.weight", optional=True):  # tied fallback, go/model.go:195-203
            raw, info = gf.get_tensor("token_embd.weight")
            raw = np.ascontiguousarray(raw)
            self._keep.append(raw)
            self.L.nlo_model_set_tensor(self.h, SLOTS["output"], -1, info.type, _p(raw), info.n_elements)
        for i in range(m.num_layers):
            for s in ("attn_norm", "ffn_norm", "attn_q", "attn_k", "attn_v", "attn_output", "ffn_gate", "ffn_up", "ffn_down"):
                put(s, i, f"blk.{i}.{s}.weight")
            for s in ("attn_q", "attn_k", "attn_v", "attn_output"):
                put(s + ".bias", i, f"blk.{i}.{s}.bias", optional=True)

    def set_gamma(self, rows: np.ndarray, token_to_row: np.ndarray):
        rows = np.ascontiguousarray(rows, dtype=np.float32)
        self._gamma = np.ascontiguousarray(token_to_row, dtype=np.int32)
        self.L.nlo_model_set_gamma(self.h, _p(rows), rows.shape[0], _p(self._gamma))

    def forward(self, token: int, pos: int) -> np.ndarray:
        if self.L.nlo_forward(self.h, int(token), int(pos)) != 0:
            raise IndexError(f"token {token} / pos {pos} out of range")
        return self.logits()

    def logits(self) -> np.ndarray:
        return np.ctypeslib.as_array(self.L.nlo_logits(self.h), shape=(self.vocab,))

    def reset(self):
        self.L.nlo_reset(self.h)

    def key_cache(self):
        n = self.cfg.n_layers * self.seq_len * self.cfg.n_kv_heads * self.cfg.head_dim
        return np.ctypeslib.as_array(self.L.nlo_key_cache(self.h), shape=(n,))

    def value_cache(self):
        n = self.cfg.n_layers * self.seq_len * self.cfg.n_kv_heads * self.cfg.head_dim
        return np.ctypeslib.as_array(self.L.nlo_value_cache(self.h), shape=(n,))

    def rope_tables(self):
        n = self.seq_len * (self.cfg.head_dim // 2)
        return (np.ctypeslib.as_array(self.L.nlo_rope_cos(self.h), shape=(n,)).copy(),
                np.ctypeslib.as_array(self.L.nlo_rope_sin(self.h), shape=(n,)).copy())

    def generate_greedy(self, prompt, n_new: int, eos_id: int = -1):
        """Engine.Generate with --temp 0 --rep-penalty 1.0 (go/main.go:152-230). returns (tokens, margins)."""
        prompt = np.ascontiguousarray(prompt, dtype=np.int32)
        out = np.zeros(n_new, dtype=np.int32)
        mg = np.zeros(n_new, dtype=np.float32)
        n = self.L.nlo_generate_greedy(self.h, _p(prompt), prompt.size, n_new, eos_id, _p(out), _p(mg))
        if n < 0:
            raise IndexError("token/pos out of range")
        return out[:n], mg[:n]

    def close(self):
        if self.h:
            self.L.nlo_model_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
