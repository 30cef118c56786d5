"""Host logic of the generation loop / samplers (go/main.go:152-408 mirror), driven on the CPU through the oracle model."""
import os
from types import SimpleNamespace

import numpy as np
import pytest

from nanollama_b200 import gguf as G
from nanollama_b200.engine import Engine, GenParams, argmax, estimate_params
from oracle import oracle as O


class OracleAdapter:
    """Gives the oracle the (*LlamaModel) surface the engine expects: forward / reset / state.logits / config."""

    def __init__(self, gf):
        self.o = O.OracleModel(gf)
        m = gf.meta
        self.config = SimpleNamespace(vocab_size=m.vocab_size, seq_len=min(m.seq_len, 2048), embed_dim=m.embed_dim, num_heads=m.num_heads,
                                      num_kv_heads=m.num_kv_heads, head_dim=m.head_dim, interm_size=m.interm_size, num_layers=m.num_layers)
        self.state = SimpleNamespace(logits=self.o.logits(), pos=0)

    def forward(self, t, p):
        self.o.forward(t, p)

    def reset(self):
        self.o.reset()


@pytest.fixture(scope="module")
def model(golden_dir):
    return OracleAdapter(G.load_gguf(os.path.join(golden_dir, "tiny_gqa_q8_0.gguf")))


@pytest.fixture(scope="module")
def prompt(golden_dir):
    return np.load(os.path.join(golden_dir, "golden_logits.npz"))["tokens"][:8]


def test_greedy_matches_reference_loop(model, prompt):
    e = Engine(model, eos_id=-1, rep_penalty=1.0)
    got = e.generate_tokens(prompt, GenParams(max_tokens=30, temperature=0.0))
    exp, _ = model.o.generate_greedy(prompt, 30)
    assert got == list(exp)


def test_eos_stops_and_is_not_emitted(model, prompt):
    exp, _ = model.o.generate_greedy(prompt, 30)
    k = next(i for i in range(2, 30) if exp[i] not in exp[:i])   # first occurrence of a token, used as the stand-in EOS
    e = Engine(model, eos_id=int(exp[k]), rep_penalty=1.0)
    got = e.generate_tokens(prompt, GenParams(max_tokens=30, temperature=0.0))
    assert got == list(exp[:k])


def test_repetition_penalty_applies_even_when_greedy(model, prompt):
    # go/main.go:177-187 mutates the logits before sampling regardless of temperature
    a = Engine(model, eos_id=-1, rep_penalty=1.0).generate_tokens(prompt, GenParams(max_tokens=40, temperature=0.0))
    b = Engine(model, eos_id=-1, rep_penalty=5.0, rep_window=64).generate_tokens(prompt, GenParams(max_tokens=40, temperature=0.0))
    assert a != b and len(set(b)) >= len(set(a))


def test_context_end_stops_generation(model, prompt):
    got = Engine(model, eos_id=-1, rep_penalty=1.0).generate_tokens(prompt, GenParams(max_tokens=500, temperature=0.0))
    assert len(got) == 64 - 8  # seq_len 64: Forward stops at pos == seq_len (go/main.go:216-218)


def test_samplers_are_seeded_and_stay_in_support(model, prompt):
    e1 = Engine(model, eos_id=-1, seed=3)
    e2 = Engine(model, eos_id=-1, seed=3)
    p = GenParams(max_tokens=12, temperature=0.8, top_p=1.0, top_k=5)
    assert e1.generate_tokens(prompt, p) == e2.generate_tokens(prompt, p)
    model.reset()
    for pos, t in enumerate(prompt):
        model.forward(int(t), pos)
    top5 = set(np.argsort(-model.state.logits)[:5].tolist())
    for _ in range(50):
        assert e1.sample_top_k(0.8, 5) in top5
        assert 0 <= e1.sample_top_p(0.8, 0.9) < 256
    assert e1.sample_top_k(0.0, 5) == e1.sample_top_p(0.0, 0.9) == argmax(model.state.logits, 256)


def test_argmax_first_maximum():
    assert argmax(np.array([1.0, 3.0, 3.0, 2.0], np.float32), 4) == 1


def test_estimate_params(model):
    assert estimate_params(model.config) == 2 * 256 * 128 + 2 * (128 * 128 + 2 * 128 * 64 + 128 * 128 + 3 * 128 * 512 + 256) + 128


class DeviceSamplingStub(OracleAdapter):
    """The two calls Engine(device_sampling=True) makes, restated on the host: forward_device = Forward without a logits copy,
    sample = what nl_sample does (penalty in place over the window, then the host samplers fed the given random number)."""

    def forward_device(self, t, p):
        self.forward(t, p)

    def sample(self, temperature, top_k, top_p, rep_penalty, recent, u):
        lg = self.state.logits
        if rep_penalty > 1.0:
            for tok in recent:
                if 0 <= tok < self.config.vocab_size:
                    lg[tok] = np.float32(lg[tok] / np.float32(rep_penalty)) if lg[tok] > 0 else np.float32(lg[tok] * np.float32(rep_penalty))
        helper = Engine(self, seed=0)
        helper.rng = SimpleNamespace(random=lambda: u)
        return helper.sample_top_p(temperature, top_p) if top_p < 1.0 else helper.sample_top_k(temperature, top_k)


@pytest.mark.parametrize("params", [GenParams(max_tokens=40, temperature=0.8, top_p=0.9, top_k=50), GenParams(max_tokens=40, temperature=1.1, top_p=1.0, top_k=7),
                                    GenParams(max_tokens=20, temperature=0.0, top_p=0.9, top_k=50)])
def test_device_sampling_loop_draws_and_windows_like_the_host_loop(golden_dir, prompt, params):
    """Engine(device_sampling=True) against the host loop, same seed: one random number per sampled token (none when greedy), the same
    repetition window, EOS and context-end handling -- so the two produce the same tokens when `sample` is exact."""
    gf = G.load_gguf(os.path.join(golden_dir, "tiny_gqa_q8_0.gguf"))
    host = Engine(OracleAdapter(gf), eos_id=-1, rep_penalty=1.15, rep_window=8, seed=11).generate_tokens(prompt, params)
    dev = Engine(DeviceSamplingStub(gf), eos_id=-1, rep_penalty=1.15, rep_window=8, seed=11, device_sampling=True).generate_tokens(prompt, params)
    assert host == dev and len(host) == params.max_tokens
    eos = host[5]
    k = host.index(eos)
    host2 = Engine(OracleAdapter(gf), eos_id=eos, rep_penalty=1.15, rep_window=8, seed=11).generate_tokens(prompt, params)
    dev2 = Engine(DeviceSamplingStub(gf), eos_id=eos, rep_penalty=1.15, rep_window=8, seed=11, device_sampling=True).generate_tokens(prompt, params)
    assert host2 == dev2 == host[:k]
