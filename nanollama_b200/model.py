"""Host-side mirror of go/model.go over the CUDA backend: same names, argument meaning and error behaviour.

    gf = load_gguf(path)                       # go/gguf.go:289   LoadGGUF
    m  = load_llama_model(gf)                  # go/model.go:121  LoadLlamaModel (weights go to HBM here)
    m.forward(token, pos)                      # go/model.go:490  (*LlamaModel).Forward — result in m.state.logits
    m.reset()                                  # go/model.go:623  (*LlamaModel).Reset

All arithmetic happens in libnanollama_cuda.so; this file only moves bytes and mirrors the Go types.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from . import capi
from .gguf import GGUFFile

_LAYER_SLOTS = ("attn_norm", "ffn_norm", "attn_q", "attn_k", "attn_v", "attn_output", "ffn_gate", "ffn_up", "ffn_down")
_BIAS_SLOTS = ("attn_q", "attn_k", "attn_v", "attn_output")


@dataclass
class LlamaConfig:
    """go/model.go:27-42."""
    num_layers: int
    embed_dim: int
    num_heads: int
    num_kv_heads: int
    head_dim: int
    vocab_size: int
    seq_len: int
    interm_size: int
    rms_norm_eps: float
    rope_theta: float
    qk_norm: bool = False
    rope_conjugate: bool = False


class LlamaState:
    """The part of go/model.go:93-118 the engine's callers touch: Logits (read AND mutated by the samplers,
    go/main.go:174-187) and Pos.  Everything else lives on the device."""

    def __init__(self, vocab: int, batch: int):
        self.logits = np.zeros(vocab, dtype=np.float32)
        self.batch_logits = np.zeros((batch, vocab), dtype=np.float32) if batch > 1 else None
        self.pos = 0


class LlamaModel:
    """go/model.go:19-24.  ``gamma`` mirrors the Gamma field (None = no gamma)."""

    def __init__(self, handle, config: LlamaConfig, max_batch: int, has_bias: bool):
        self._h = handle
        self.config = config
        self.state = LlamaState(config.vocab_size, max_batch)
        self.gamma = None
        self.max_batch = max_batch
        self.has_bias = has_bias
        self.launches_last_prefill = 0

    # -- (*LlamaModel).Forward: no return value; OOB token/pos is a panic in Go -> IndexError here
    def forward(self, token: int, pos: int) -> None:
        rc = capi.lib().nl_forward(self._h, int(token), int(pos), capi.ptr(self.state.logits))
        if rc == capi.NL_ERR_INVALID:
            raise IndexError(capi.lib().nl_last_error().decode())
        capi.check(rc)

    def forward_device(self, token: int, pos: int) -> None:
        """Forward whose logits stay on the device (nl_forward with logits_out = NULL): the companion of ``sample``."""
        rc = capi.lib().nl_forward(self._h, int(token), int(pos), None)
        if rc == capi.NL_ERR_INVALID:
            raise IndexError(capi.lib().nl_last_error().decode())
        capi.check(rc)

    def sample(self, temperature: float, top_k: int, top_p: float, rep_penalty: float, recent: Sequence[int], u: float) -> int:
        """One sampling step of Engine.Generate on the device-resident logits (go/main.go:177-197, :294-398): nl_sample.
        ``u`` is the uniform number the reference's rng.Float32() would deliver at this step."""
        r = np.ascontiguousarray(recent, dtype=np.int32)
        tok = C.c_int32(0)
        rc = capi.lib().nl_sample(self._h, float(temperature), int(top_k), float(top_p), float(rep_penalty), capi.ptr(r) if r.size else None, int(r.size),
                                  float(u), C.byref(tok))
        capi.check(rc)
        return int(tok.value)

    def get_logits(self) -> np.ndarray:
        """State.Logits read-back of the device-resident logits (nl_get_logits)."""
        capi.check(capi.lib().nl_get_logits(self._h, capi.ptr(self.state.logits)))
        return self.state.logits

    def reset(self) -> None:
        capi.check(capi.lib().nl_reset(self._h))
        self.state.pos = 0

    # -- backend extensions (no counterpart in the reference) --
    def forward_batch(self, tokens: Sequence[int], pos: Sequence[int]) -> np.ndarray:
        t = np.ascontiguousarray(tokens, dtype=np.int32)
        p = np.ascontiguousarray(pos, dtype=np.int32)
        out = np.empty((t.size, self.config.vocab_size), dtype=np.float32)
        rc = capi.lib().nl_forward_batch(self._h, t.size, capi.ptr(t), capi.ptr(p), capi.ptr(out))
        if rc == capi.NL_ERR_INVALID:
            raise IndexError(capi.lib().nl_last_error().decode())
        capi.check(rc)
        return out

    def prefill(self, tokens: Sequence[int], pos0: int = 0) -> None:
        t = np.ascontiguousarray(tokens, dtype=np.int32)
        rc = capi.lib().nl_prefill(self._h, capi.ptr(t), t.size, int(pos0), capi.ptr(self.state.logits))
        if rc == capi.NL_ERR_INVALID:
            raise IndexError(capi.lib().nl_last_error().decode())
        capi.check(rc)

    def generate_greedy(self, prompt: Sequence[int], n_new: int, eos_id: int = -1) -> np.ndarray:
        p = np.ascontiguousarray(prompt, dtype=np.int32)
        out = np.zeros(max(n_new, 1), dtype=np.int32)
        n = C.c_int32(0)
        capi.check(capi.lib().nl_generate_greedy(self._h, capi.ptr(p), p.size, int(n_new), int(eos_id), capi.ptr(out), C.byref(n)))
        return out[: n.value]

    def set_gamma(self, rows: Optional[np.ndarray], token_to_row: Optional[np.ndarray]) -> None:
        if rows is None:
            capi.check(capi.lib().nl_set_gamma(self._h, None, 0, None))
            self.gamma = None
            return
        rows = np.ascontiguousarray(rows, dtype=np.float32)
        mp = np.ascontiguousarray(token_to_row, dtype=np.int32)
        if rows.ndim != 2 or rows.shape[1] != self.config.embed_dim or mp.size != self.config.vocab_size:
            raise ValueError("gamma shape mismatch")
        capi.check(capi.lib().nl_set_gamma(self._h, capi.ptr(rows), rows.shape[0], capi.ptr(mp)))
        self.gamma = (rows, mp)

    def bench_decode(self, token: int, pos0: int, n_steps: int) -> float:
        ms = C.c_float(0)
        capi.check(capi.lib().nl_bench_decode(self._h, int(token), int(pos0), int(n_steps), C.byref(ms)))
        return ms.value

    def bench_prefill(self, tokens: Sequence[int], pos0: int = 0) -> float:
        """CUDA-event time (ms) of the prefill's kernels, prompt already on the device (nl_bench_prefill)."""
        t = np.ascontiguousarray(tokens, dtype=np.int32)
        ms, nl = C.c_float(0), C.c_int32(0)
        capi.check(capi.lib().nl_bench_prefill(self._h, capi.ptr(t), t.size, int(pos0), C.byref(ms), C.byref(nl)))
        self.launches_last_prefill = int(nl.value)
        return ms.value

    @property
    def launches_per_token(self) -> int:
        return capi.lib().nl_launches_per_token(self._h)

    @property
    def decode_path(self) -> str:
        return capi.lib().nl_decode_path(self._h).decode()

    @property
    def weight_bytes(self) -> int:
        return capi.lib().nl_weight_bytes(self._h)

    def close(self) -> None:
        if self._h:
            capi.lib().nl_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def load_llama_model(gf: GGUFFile, device: int = 0, max_batch: int = 1, verbose: bool = False,
                     rope_conjugate: Optional[bool] = None, qk_norm: Optional[bool] = None,
                     tp_rank: int = 0, tp_size: int = 1, exchange=None) -> LlamaModel:
    """go/model.go:121-174 + loadWeights :177-265.  Errors are wrapped like the Go ones ("load weights: layer 3 attn_q: ...").

    tp_size > 1: this process is rank `tp_rank` of a tensor-parallel group (one process per GPU); every rank passes the FULL
    tensors and the library keeps its shard.  `exchange(handle_bytes) -> [handle of rank 0, 1, ...]` all-gathers the 64-byte
    CUDA IPC handles of the ranks' NVLink windows (nanollama_b200.tp.exchange_handles_torch by default)."""
    m = gf.meta
    L = capi.lib()
    cfg = capi.NlConfig(m.num_layers, m.embed_dim, m.num_heads, m.num_kv_heads, m.head_dim, m.vocab_size, m.seq_len, m.interm_size,
                        m.rms_norm_eps, m.rope_theta, int(m.qk_norm if qk_norm is None else qk_norm),
                        int(m.rope_conjugate if rope_conjugate is None else rope_conjugate), device, tp_rank, tp_size, max_batch)
    h = C.c_void_p()
    capi.check(L.nl_create(C.byref(cfg), C.byref(h)))
    try:
        def put(slot: str, layer: int, name: str, optional: bool = False) -> bool:
            try:
                raw, info = gf.get_tensor(name)
            except KeyError:
                if optional:
                    return False
                raise
            raw = np.ascontiguousarray(raw)
            rows, cols = info.rows_cols
            rc = L.nl_upload_tensor(h, capi.SLOTS[slot], layer, info.type, rows, cols, capi.ptr(raw), raw.size)
            if rc != capi.NL_OK:
                raise capi.NlError(rc, f"{name}: " + L.nl_last_error().decode())
            return True

        try:
            put("token_embd", -1, "token_embd.weight")
            put("output_norm", -1, "output_norm.weight")
            if not put("output", -1, "output.weight", optional=True) and verbose:
                print("[model] output.weight not found, using tied embeddings")  # go/model.go:195-203
            has_bias = False
            for i in range(m.num_layers):
                for s in _LAYER_SLOTS:
                    put(s, i, f"blk.{i}.{s}.weight")
                for s in _BIAS_SLOTS:
                    has_bias |= put(s + ".bias", i, f"blk.{i}.{s}.bias", optional=True)
        except (KeyError, capi.NlError) as e:
            raise RuntimeError(f"load weights: {e}") from e
        capi.check(L.nl_finalize(h))
        if tp_size > 1:
            if exchange is None:
                from .tp import exchange_handles_torch as exchange
            mine = C.create_string_buffer(64)
            capi.check(L.nl_tp_export_handle(h, mine))
            handles = exchange(mine.raw)
            if len(handles) != tp_size or any(len(x) != 64 for x in handles):
                raise RuntimeError("tensor-parallel handle exchange returned a malformed list")
            blob = C.create_string_buffer(b"".join(handles), 64 * tp_size)
            capi.check(L.nl_tp_import_handles(h, blob, tp_size))
        out = capi.NlConfig()
        capi.check(L.nl_get_config(h, C.byref(out)))
    except Exception:
        L.nl_destroy(h)
        raise
    config = LlamaConfig(out.n_layers, out.embed_dim, out.n_heads, out.n_kv_heads, out.head_dim, out.vocab_size, out.seq_len, out.interm_size,
                         out.rms_norm_eps, out.rope_theta, bool(out.qk_norm), bool(out.rope_conjugate))
    if verbose:
        if m.seq_len > 2048:
            print(f"[model] capping seq_len from {m.seq_len} to 2048")
        print(f"[model] loaded: {config.num_layers} layers, {config.embed_dim} dim, {config.num_heads} heads, "
              f"{config.num_kv_heads} kv_heads, {config.vocab_size} vocab, bias={has_bias}")
    return LlamaModel(h, config, max_batch, has_bias)


# ---- operator-level hooks (matmulDispatch / Dequant*, go/model.go:361-386, go/quant.go) ----
def dequant(ggml_type: int, raw: np.ndarray, n: int) -> np.ndarray:
    raw = np.ascontiguousarray(raw, dtype=np.uint8)
    out = np.empty(n, dtype=np.float32)
    capi.check(capi.lib().nl_dequant(int(ggml_type), capi.ptr(raw), n, capi.ptr(out)))
    return out


def matmul_dispatch(raw: np.ndarray, ggml_type: int, x: np.ndarray, rows: int, cols: int) -> np.ndarray:
    """x: [cols] or [batch, cols] fp32 -> [rows] or [batch, rows]."""
    raw = np.ascontiguousarray(raw, dtype=np.uint8)
    x2 = np.ascontiguousarray(x, dtype=np.float32).reshape(-1, cols)
    out = np.empty((x2.shape[0], rows), dtype=np.float32)
    capi.check(capi.lib().nl_matmul(int(ggml_type), capi.ptr(raw), rows, cols, capi.ptr(x2), x2.shape[0], capi.ptr(out)))
    return out[0] if np.ndim(x) == 1 else out


class DeviceMatrix:
    """A weight matrix resident in HBM (for repeated GEMVs and the microbench)."""

    def __init__(self, raw: np.ndarray, ggml_type: int, rows: int, cols: int, device: int = 0):
        raw = np.ascontiguousarray(raw, dtype=np.uint8)
        self._h = C.c_void_p()
        self.rows, self.cols, self.type = rows, cols, ggml_type
        capi.check(capi.lib().nl_matrix_create(int(ggml_type), capi.ptr(raw), rows, cols, device, C.byref(self._h)))

    def matmul(self, x: np.ndarray) -> np.ndarray:
        x2 = np.ascontiguousarray(x, dtype=np.float32).reshape(-1, self.cols)
        out = np.empty((x2.shape[0], self.rows), dtype=np.float32)
        capi.check(capi.lib().nl_matrix_matmul(self._h, capi.ptr(x2), x2.shape[0], capi.ptr(out)))
        return out[0] if np.ndim(x) == 1 else out

    def bench(self, batch: int = 1, n_copies: int = 1, warmup: int = 5, iters: int = 50) -> float:
        ms = C.c_float(0)
        capi.check(capi.lib().nl_matrix_bench(self._h, batch, n_copies, warmup, iters, C.byref(ms)))
        return ms.value

    def close(self):
        if self._h:
            capi.lib().nl_matrix_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
