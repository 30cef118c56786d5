"""Tier shapes of the reference (nanollama/llama.py:40-51, FFN width :178-179, vocab tiers README.md:49-54) and
deterministic random-init GGUF models of those shapes, generated DIRECTLY as quantized blocks:
the reference's pure-Python producers run at 0.1-0.6 M elem/s and cannot emit anything above nano.

``SyntheticGGUF`` quacks like ``gguf.GGUFFile`` (``.meta``, ``.tensors``, ``.get_tensor``) but makes each tensor on demand
from a per-tensor seeded generator, so the 4.2 GB `big` tier never has to touch the disk; ``write_gguf`` streams the same
bytes into a real file (with an embedded synthetic token list, which is where the Go engine takes VocabSize from).
"""
from __future__ import annotations

import zlib
from typing import Dict, Tuple

import numpy as np

from . import gguf as G

#            L   dim   H  KV  vocab
TIERS: Dict[str, Tuple[int, int, int, int, int]] = {
    "nano":   (13,  576,  9,  9, 32000),
    "micro":  (16,  640, 10, 10, 32000),
    "mini":   (20,  768, 12,  3, 32000),
    "small":  (24, 1024, 16,  4, 32000),
    "goldie": (28, 1536, 24,  6, 48000),
    "medium": (32, 2048, 32,  8, 64000),
    "large":  (36, 3072, 48, 12, 96000),
    "big":    (40, 4096, 64, 16, 96000),
}


def interm_size(dim: int, multiple_of: int = 256) -> int:
    """nanollama/llama.py:178-179 == scripts/export_gguf.py:348-351."""
    hidden = int(2 * (4 * dim) / 3)
    return multiple_of * ((hidden + multiple_of - 1) // multiple_of)


def tier_meta(tier: str, seq_len: int = 2048, vocab: int | None = None, layers: int | None = None) -> G.GGUFMetadata:
    L, dim, H, KV, V = TIERS[tier]
    m = G.GGUFMetadata(num_layers=layers or L, embed_dim=dim, num_heads=H, num_kv_heads=KV, head_dim=dim // H,
                       vocab_size=vocab or V, seq_len=seq_len, interm_size=interm_size(dim))
    return m


def matmul_params(m: G.GGUFMetadata) -> int:
    """weight elements touched per decoded token (SURVEY.md §8 tier table)."""
    kvd = m.num_kv_heads * m.head_dim
    return m.num_layers * (2 * m.embed_dim ** 2 + 2 * kvd * m.embed_dim + 3 * m.embed_dim * m.interm_size) + m.vocab_size * m.embed_dim


BPE = {G.GGML_Q4_0: 18 / 32, G.GGML_Q8_0: 34 / 32, G.GGML_F16: 2.0, G.GGML_F32: 4.0}


def decode_bytes_per_token(m: G.GGUFMetadata, ggml_type: int, pos: int, kv_bytes: int = 4) -> int:
    """Algorithmic HBM bytes of one bs=1 decode step at position `pos`."""
    kvd = m.num_kv_heads * m.head_dim
    bpe = BPE[ggml_type]
    w = bpe * matmul_params(m) + bpe * m.embed_dim + 4 * (2 * m.num_layers + 1) * m.embed_dim
    kv = kv_bytes * 2 * m.num_layers * kvd * (pos + 1) + kv_bytes * 2 * m.num_layers * kvd
    return int(w + kv)


class SyntheticGGUF:
    """Random-init model of a tier in GGUF tensor encoding, generated on demand and reproducibly."""

    def __init__(self, tier: str, ggml_type: int, seed: int = 0, seq_len: int = 2048, vocab: int | None = None, layers: int | None = None):
        if ggml_type not in (G.GGML_Q4_0, G.GGML_Q8_0, G.GGML_F16):
            raise ValueError("synthetic tiers are Q4_0, Q8_0 or F16")
        self.tier, self.type, self.seed = tier, ggml_type, seed
        self.meta = tier_meta(tier, seq_len, vocab, layers)
        m = self.meta
        kvd = m.num_kv_heads * m.head_dim
        self.tensors: Dict[str, G.GGUFTensorInfo] = {}
        off = 0

        def add(name, shape, t):
            nonlocal off
            off = (off + 31) // 32 * 32
            info = G.GGUFTensorInfo(name, len(shape), tuple(reversed(shape)), t, off)
            self.tensors[name] = info
            off += G.tensor_bytes(info)

        add("token_embd.weight", (m.vocab_size, m.embed_dim), ggml_type)
        add("output_norm.weight", (m.embed_dim,), G.GGML_F32)
        add("output.weight", (m.vocab_size, m.embed_dim), ggml_type)
        for i in range(m.num_layers):
            p = f"blk.{i}."
            add(p + "attn_norm.weight", (m.embed_dim,), G.GGML_F32)
            add(p + "ffn_norm.weight", (m.embed_dim,), G.GGML_F32)
            add(p + "attn_q.weight", (m.embed_dim, m.embed_dim), ggml_type)
            add(p + "attn_k.weight", (kvd, m.embed_dim), ggml_type)
            add(p + "attn_v.weight", (kvd, m.embed_dim), ggml_type)
            add(p + "attn_output.weight", (m.embed_dim, m.embed_dim), ggml_type)
            add(p + "ffn_gate.weight", (m.interm_size, m.embed_dim), ggml_type)
            add(p + "ffn_up.weight", (m.interm_size, m.embed_dim), ggml_type)
            add(p + "ffn_down.weight", (m.embed_dim, m.interm_size), ggml_type)
        self.total_bytes = off

    def _rng(self, name: str) -> np.random.Generator:
        return np.random.Generator(np.random.PCG64([self.seed, zlib.crc32(name.encode())]))

    def get_tensor(self, name: str):
        info = self.tensors.get(name)
        if info is None:
            raise KeyError(f"tensor not found: {name}")
        rng = self._rng(name)
        n = info.n_elements
        if info.type == G.GGML_F32:  # norm weights 1 + 0.1 N(0,1)
            return (1.0 + 0.1 * rng.standard_normal(n)).astype(np.float32).view(np.uint8), info
        cols = info.dims[0]
        # token embedding rows feed the residual stream directly: unit variance; projections: variance-preserving
        target_std = 1.0 if name == "token_embd.weight" else 1.0 / np.sqrt(cols)
        if info.type == G.GGML_F16:
            return (target_std * rng.standard_normal(n, dtype=np.float32)).astype(np.float16).view(np.uint8), info
        nb = n // 32
        u = rng.uniform(0.5, 1.5, size=nb).astype(np.float32)
        if info.type == G.GGML_Q4_0:   # codes uniform over 0..15 -> std(code-8) = 4.61
            out = np.empty((nb, 18), dtype=np.uint8)
            out[:, :2] = (u * np.float32(target_std / 4.61)).astype(np.float16).view(np.uint8).reshape(nb, 2)
            out[:, 2:] = rng.integers(0, 256, size=(nb, 16), dtype=np.uint8)
        else:                          # int8 uniform over -128..127 -> std 73.9
            out = np.empty((nb, 34), dtype=np.uint8)
            out[:, :2] = (u * np.float32(target_std / 73.9)).astype(np.float16).view(np.uint8).reshape(nb, 2)
            out[:, 2:] = rng.integers(0, 256, size=(nb, 32), dtype=np.uint8)
        return out.reshape(-1), info

    def find_tensor(self, substr: str):
        for n, i in self.tensors.items():
            if substr in n:
                return i
        return None

    def write_gguf(self, path: str) -> None:
        """Stream this model into a real GGUF v3 file with the KV set of scripts/export_gguf.py:520-537 and a token list."""
        m = self.meta
        w = G.GGUFWriter(path)
        w.add_string("general.architecture", "llama")
        w.add_string("general.name", f"nanollama-{self.tier}-synthetic")
        w.add_uint32("llama.block_count", m.num_layers)
        w.add_uint32("llama.embedding_length", m.embed_dim)
        w.add_uint32("llama.attention.head_count", m.num_heads)
        w.add_uint32("llama.attention.head_count_kv", m.num_kv_heads)
        w.add_uint32("llama.attention.key_length", m.head_dim)
        w.add_uint32("llama.attention.value_length", m.head_dim)
        w.add_uint32("llama.feed_forward_length", m.interm_size)
        w.add_uint32("llama.context_length", m.seq_len)
        w.add_float32("llama.attention.layer_norm_rms_epsilon", m.rms_norm_eps)
        w.add_float32("llama.rope.freq_base", m.rope_theta)
        w.add_uint32("llama.vocab_size", m.vocab_size)
        w.add_bool("nanollama.qk_norm", m.qk_norm)
        w.add_bool("nanollama.rope_conjugate", m.rope_conjugate)
        w.add_string("tokenizer.ggml.model", "llama")
        w.add_array("tokenizer.ggml.tokens", G.T_STRING, ["<unk>", "<s>", "</s>"] + [f"<t{i}>" for i in range(3, m.vocab_size)])
        w.add_uint32("tokenizer.ggml.bos_token_id", 1)
        w.add_uint32("tokenizer.ggml.eos_token_id", 2)
        for name, info in self.tensors.items():
            w.add_tensor_raw(name, (lambda n=name: self.get_tensor(n)[0]), info.type, tuple(reversed(info.dims)))
        w.write()
