"""Scheduling logic of the continuous batcher (nanollama_b200/batcher.py; replaces the global mutex of go/serve.go:56,106-108) on a
fake model whose logits depend only on the sequence's own token history: every request must get exactly the tokens the
single-sequence loop (engine.Engine.generate_tokens = GenerateQuiet, go/main.go:233-291) produces for the same seed.  CPU only."""
import hashlib
import threading

import numpy as np
import pytest

from nanollama_b200.batcher import ContinuousBatcher
from nanollama_b200.engine import Engine, GenParams


class _Cfg:
    def __init__(self, vocab, seq_len):
        self.vocab_size, self.seq_len = vocab, seq_len


class _State:
    def __init__(self, vocab):
        self.logits = np.zeros(vocab, np.float32)


class FakeModel:
    """forward / forward_batch / reset with per-slot histories; logits = a hash of the slot's tokens so far (position-checked)."""

    def __init__(self, vocab=97, seq_len=48, max_batch=4):
        self.config, self.state, self.max_batch = _Cfg(vocab, seq_len), _State(vocab), max_batch
        self.hist = [[] for _ in range(max_batch)]
        self.batch_sizes = []

    def _logits(self, b, token, pos):
        assert 0 <= pos < self.config.seq_len and pos <= len(self.hist[b])
        self.hist[b] = self.hist[b][:pos] + [int(token)]
        h = hashlib.sha256(np.asarray(self.hist[b], np.int64).tobytes()).digest()
        rng = np.random.default_rng(int.from_bytes(h[:8], "little"))
        return (3.0 * rng.standard_normal(self.config.vocab_size)).astype(np.float32)

    def reset(self):
        self.hist[0] = []

    def forward(self, token, pos):
        self.batch_sizes.append(1)
        self.state.logits[:] = self._logits(0, token, pos)

    def forward_batch(self, tokens, pos):
        self.batch_sizes.append(len(tokens))
        return np.stack([self._logits(b, t, p) for b, (t, p) in enumerate(zip(tokens, pos))])


def _alone(prompt, p, seed, **kw):
    return Engine(FakeModel(max_batch=1), seed=seed, **kw).generate_tokens(prompt, p)


REQS = [([1, 5, 9], GenParams(max_tokens=12, temperature=0.8, top_p=0.9, top_k=50), 1),
        ([1, 7], GenParams(max_tokens=30, temperature=1.1, top_p=1.0, top_k=8), 2),
        ([1, 2, 3, 4, 5, 6, 7, 8, 9, 10], GenParams(max_tokens=5, temperature=0.0, top_p=0.9, top_k=50), 3),
        ([4] * 60, GenParams(max_tokens=9, temperature=0.7, top_p=0.5, top_k=50), 4),        # prompt longer than the context
        ([1, 3], GenParams(max_tokens=100, temperature=0.9, top_p=0.95, top_k=50), 5),       # runs into the end of the context
        ([2, 2, 2], GenParams(max_tokens=0, temperature=0.8, top_p=0.9, top_k=50), 6),
        ([1, 11, 12], GenParams(max_tokens=20, temperature=0.8, top_p=0.9, top_k=50), 7)]


@pytest.mark.parametrize("max_batch", [1, 2, 4])
def test_every_request_gets_its_single_sequence_stream(max_batch):
    m = FakeModel(max_batch=max_batch)
    b = ContinuousBatcher(m)
    tickets = [b.submit(pr, p, seed=s) for pr, p, s in REQS]
    b.run_until_idle()
    for (pr, p, s), t in zip(REQS, tickets):
        assert t.result(0) == _alone(pr, p, s), (pr, p)
    assert max(m.batch_sizes) <= max_batch
    if max_batch > 1:
        assert max(m.batch_sizes) == max_batch and b.rows_run / b.steps_run > 1.5   # sequences really shared steps
    assert all(s is None for s in b.slots) and not b.waiting


def test_eos_and_late_arrivals_and_callbacks():
    m = FakeModel(max_batch=3)
    b = ContinuousBatcher(m, eos_id=13, rep_penalty=1.3, rep_window=4)
    seen = []
    t1 = b.submit([1, 5], GenParams(max_tokens=40, temperature=0.9, top_p=0.9, top_k=50), seed=11, on_token=seen.append)
    for _ in range(5):
        b.step()
    t2 = b.submit([1, 6, 7], GenParams(max_tokens=25, temperature=0.9, top_p=1.0, top_k=5), seed=12)   # joins while t1 is in flight
    b.run_until_idle()
    kw = dict(eos_id=13, rep_penalty=1.3, rep_window=4)
    e1 = _alone([1, 5], GenParams(max_tokens=40, temperature=0.9, top_p=0.9, top_k=50), 11, **kw)
    assert t1.result(0) == e1 and seen == e1
    assert t2.result(0) == _alone([1, 6, 7], GenParams(max_tokens=25, temperature=0.9, top_p=1.0, top_k=5), 12, **kw)
    assert 13 not in t1.tokens and 13 not in t2.tokens
    assert b.submit([], GenParams(), seed=0).error is not None


def test_worker_thread_serves_concurrent_submitters():
    m = FakeModel(max_batch=4)
    b = ContinuousBatcher(m)
    b.start()
    out = {}

    def client(i):
        pr, p, s = REQS[i % len(REQS)]
        out[i] = (b.submit(pr, p, seed=s + i).result(30), _alone(pr, p, s + i))

    th = [threading.Thread(target=client, args=(i,)) for i in range(10)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    b.stop()
    assert len(out) == 10 and all(got == exp for got, exp in out.values())


def test_a_failing_forward_fails_the_requests_not_the_scheduler():
    m = FakeModel(max_batch=2)
    b = ContinuousBatcher(m)
    t = b.submit([96 + 5], GenParams(max_tokens=3), seed=0)   # token outside the fake vocabulary range is fine; break the model instead

    def boom(*a):
        raise IndexError("token out of range")
    m.forward = boom
    b.run_until_idle()
    with pytest.raises(IndexError):
        t.result(0)
    m2 = FakeModel(max_batch=2)
    b.model = m2
    t2 = b.submit([1, 2], GenParams(max_tokens=4, temperature=0.0), seed=0)
    b.run_until_idle()
    assert t2.result(0) == _alone([1, 2], GenParams(max_tokens=4, temperature=0.0), 0)
