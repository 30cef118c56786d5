"""Host-side mirror of the reference's generation loop and samplers (go/main.go:143-408) on top of any model object that has
``forward(token, pos)``, ``reset()``, ``state.logits`` and ``config`` — i.e. nanollama_b200.model.LlamaModel (CUDA backend).

The loop is host code in the reference too (it reads and mutates State.Logits between Forward calls), so it stays on the
host here; ``LlamaModel.generate_greedy`` is the all-on-device shortcut for temp<=0 / rep-penalty 1.0.
"""
from __future__ import annotations

import math
import random
from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence

import numpy as np


@dataclass
class GenParams:
    """go/main.go:135-141 (defaults of the CLI flags, main.go:29-34)."""
    max_tokens: int = 256
    temperature: float = 0.8
    top_p: float = 0.9
    top_k: int = 50


def argmax(logits: np.ndarray, n: int) -> int:
    """go/main.go:400-408: first maximum, strict '>'."""
    return int(np.argmax(logits[:n]))  # numpy returns the first occurrence of the maximum


def sample_top_k(logits: np.ndarray, vocab: int, temp: float, top_k: int, rng) -> int:
    """sampleTopK, go/main.go:294-343 (one rng draw when temp > 0)."""
    if temp <= 0:
        return argmax(logits, vocab)
    top_k = min(top_k, vocab)
    order = np.argsort(-logits[:vocab], kind="stable")[:top_k]  # insertion order of the reference == stable descending
    vals = logits[order]
    probs = np.exp(((vals - vals[0]) / np.float32(temp)).astype(np.float64)).astype(np.float32)
    total = np.float32(0)
    for p in probs:
        total = np.float32(total + p)
    r = np.float32(rng.random()) * total
    cdf = np.float32(0)
    for i, p in enumerate(probs):
        cdf = np.float32(cdf + p)
        if r <= cdf:
            return int(order[i])
    return int(order[0])


def sample_top_p(logits: np.ndarray, vocab: int, temp: float, top_p: float, rng) -> int:
    """sampleTopP, go/main.go:346-398."""
    if temp <= 0:
        return argmax(logits, vocab)
    lg = logits[:vocab]
    mx = lg.max()
    p = np.exp(((lg - mx) / np.float32(temp)).astype(np.float64)).astype(np.float32)
    s = np.float32(0)
    for v in p:  # fp32 sequential sum like the reference
        s = np.float32(s + v)
    p = p * (np.float32(1.0) / s)
    order = np.argsort(-p, kind="stable")
    cum = np.float32(0)
    for i, idx in enumerate(order):
        cum = np.float32(cum + p[idx])
        if cum >= top_p:
            r = np.float32(rng.random()) * cum
            cdf = np.float32(0)
            for j in range(i + 1):
                cdf = np.float32(cdf + p[order[j]])
                if r <= cdf:
                    return int(order[j])
            return int(order[0])
    return int(order[0])


class Engine:
    """go/main.go:143-149."""

    def __init__(self, model, eos_id: int = 2, rep_penalty: float = 1.15, rep_window: int = 64, seed: Optional[int] = None,
                 decode_token: Optional[Callable[[int], str]] = None, device_sampling: bool = False):
        # device_sampling: repetition penalty + sampleTopK / sampleTopP run on the device (nl_sample) on logits that never leave HBM;
        # the random numbers are still drawn here, one per sampled token, exactly where the host samplers draw theirs
        self.device_sampling = bool(device_sampling)
        self.model = model
        self.eos_id = eos_id
        self.rep_penalty = float(rep_penalty)
        self.rep_window = int(rep_window)
        self.rng = random.Random(seed)
        self.decode_token = decode_token or (lambda t: "")

    # -- go/main.go:294-343 / :346-398 on State.Logits
    def sample_top_k(self, temp: float, top_k: int) -> int:
        return sample_top_k(self.model.state.logits, self.model.config.vocab_size, temp, top_k, self.rng)

    def sample_top_p(self, temp: float, top_p: float) -> int:
        return sample_top_p(self.model.state.logits, self.model.config.vocab_size, temp, top_p, self.rng)

    # -- go/main.go:233-291 (GenerateQuiet); Generate (:152-230) is the same loop plus stdout streaming
    def generate_tokens(self, prompt_tokens: Sequence[int], p: GenParams, on_token: Optional[Callable[[int], None]] = None) -> List[int]:
        m = self.model
        seq_len = m.config.seq_len
        dev = self.device_sampling
        fwd = m.forward_device if dev else m.forward
        m.reset()
        pos = 0
        for tok in prompt_tokens:
            fwd(int(tok), pos)
            pos += 1
            if pos >= seq_len - 1:
                break
        out: List[int] = []
        out_bytes = 0
        recent: List[int] = []
        for _ in range(p.max_tokens):
            if out_bytes >= 8192:
                break
            if dev:
                # the host samplers draw one rng.Float32() per sampled token, and none when temperature <= 0 (main.go:299, :351)
                u = float(np.float32(self.rng.random())) if p.temperature > 0 else 0.0
                if u >= 1.0:
                    u = float(np.nextafter(np.float32(1.0), np.float32(0.0)))
                nxt = m.sample(p.temperature, p.top_k, p.top_p, self.rep_penalty, recent, u)
                recent.append(nxt)
                if len(recent) > self.rep_window:
                    recent = recent[1:]
                if nxt == self.eos_id:
                    break
                out.append(nxt)
                out_bytes += len(self.decode_token(nxt).encode("utf-8"))
                if on_token:
                    on_token(nxt)
                fwd(nxt, pos)
                pos += 1
                if pos >= seq_len:
                    break
                continue
            logits = m.state.logits
            if self.rep_penalty > 1.0 and recent:  # go/main.go:177-187 — in place, also when temp == 0
                for tok in recent:
                    if 0 <= tok < m.config.vocab_size:
                        if logits[tok] > 0:
                            logits[tok] = np.float32(logits[tok] / np.float32(self.rep_penalty))
                        else:
                            logits[tok] = np.float32(logits[tok] * np.float32(self.rep_penalty))
            nxt = self.sample_top_p(p.temperature, p.top_p) if p.top_p < 1.0 else self.sample_top_k(p.temperature, p.top_k)
            recent.append(nxt)
            if len(recent) > self.rep_window:
                recent = recent[1:]
            if nxt == self.eos_id:
                break
            out.append(nxt)
            out_bytes += len(self.decode_token(nxt).encode("utf-8"))
            if on_token:
                on_token(nxt)
            m.forward(nxt, pos)
            pos += 1
            if pos >= seq_len:
                break
        return out

    def generate(self, prompt_tokens: Sequence[int], p: GenParams) -> str:
        return "".join(self.decode_token(t) for t in self.generate_tokens(prompt_tokens, p))


def estimate_params(cfg) -> int:
    """go/main.go:411-425."""
    embed = cfg.vocab_size * cfg.embed_dim
    attn = cfg.embed_dim * (cfg.num_heads * cfg.head_dim) + 2 * cfg.embed_dim * (cfg.num_kv_heads * cfg.head_dim) + cfg.embed_dim * cfg.embed_dim
    mlp = 3 * cfg.embed_dim * cfg.interm_size
    return 2 * embed + cfg.num_layers * (attn + mlp + 2 * cfg.embed_dim) + cfg.embed_dim
