"""Pins the CPU oracle (oracle/nl_oracle.c) against fixtures produced by the reference's own code
(tests/golden/make_golden.py): reference quantizer bytes decoded by gguf-py, and logits / greedy streams of the
reference torch model.  CPU only."""
import os

import numpy as np
import pytest

from nanollama_b200 import gguf as G
from oracle import oracle as O


@pytest.fixture(scope="module")
def kat(golden_dir):
    return np.load(os.path.join(golden_dir, "dequant_kat.npz"))


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "golden_logits.npz"))


def test_half2float_all_65536(kat):
    # go/gguf.go:601-636 LUT == IEEE binary16->binary32 for every bit pattern (NaN payloads included)
    # (through memory, not a float return register, so signalling-NaN payloads are not quieted on the way)
    allh = np.arange(65536, dtype=np.uint16)
    got = O.dequant(G.GGML_F16, allh.view(np.uint8), 65536).view(np.uint32)
    assert np.array_equal(got, kat["half_all_expect"])


@pytest.mark.parametrize("key,typ", [("q4_0", G.GGML_Q4_0), ("q8_0", G.GGML_Q8_0), ("q8_0_requant", G.GGML_Q8_0), ("f16", G.GGML_F16)])
def test_dequant_bit_exact_vs_reference_bytes(kat, key, typ):
    raw = kat[key + "_bytes"]
    got = O.dequant(typ, raw, 2048).view(np.uint32)
    assert np.array_equal(got, kat[key + "_expect"])


KQ = [("q5_0", G.GGML_Q5_0), ("q4_k", G.GGML_Q4_K), ("q6_k", G.GGML_Q6_K)]


@pytest.fixture(scope="module")
def kq(golden_dir):
    return np.load(os.path.join(golden_dir, "kquant_kat.npz"))


@pytest.mark.parametrize("key,typ", KQ)
def test_kquant_dequant_bit_exact_vs_gguf_py(kq, key, typ):
    """DequantQ5_0 / DequantQ4_K / DequantQ6_K (go/quant.go:171-484) against gguf-py's independent decoder
    (tests/golden/make_kquant_kat.py): every value, bit for bit."""
    exp = kq[key + "_expect"]
    got = O.dequant(typ, kq[key + "_bytes"], exp.size).view(np.uint32)
    assert np.array_equal(got, exp)


@pytest.mark.parametrize("key,typ", KQ)
def test_kquant_matmul_vs_gguf_py_weights(kq, key, typ):
    """MatMulQ5_0 / Q4_K / Q6_K: the oracle's product against a float64 dot with the gguf-py-decoded weights."""
    w = kq[key + "_expect"].view(np.float32)
    cols = 1024
    rows = w.size // cols
    x = np.random.default_rng(7).standard_normal(cols).astype(np.float32)
    got = O.matmul(kq[key + "_bytes"], typ, x, rows, cols)
    exp = w.reshape(rows, cols).astype(np.float64) @ x.astype(np.float64)
    assert np.abs(got - exp).max() / np.abs(exp).max() < 1e-5


def test_matmul_equals_dequant_then_dot_ordering(kat):
    # MatMulQ4_0 applies the scale after the per-block dot (go/quant.go:83-90); check against an explicit restatement
    rng = np.random.default_rng(0)
    rows, cols = 8, 256
    w = rng.standard_normal((rows, cols)).astype(np.float32)
    x = rng.standard_normal(cols).astype(np.float32)
    for typ, q in ((G.GGML_Q4_0, G.quantize_q4_0(w)), (G.GGML_Q8_0, G.quantize_q8_0(w))):
        got = O.matmul(q, typ, x, rows, cols)
        bs = G.ggml_block_size(typ)
        blocks = q.reshape(rows, cols // 32, bs)
        exp = np.zeros(rows, dtype=np.float32)
        for r in range(rows):
            s = np.float32(0)
            for b in range(cols // 32):
                d = blocks[r, b, :2].view(np.float16).astype(np.float32)[0]
                dot = np.float32(0)
                if typ == G.GGML_Q4_0:
                    for j in range(16):
                        bv = int(blocks[r, b, 2 + j])
                        v0 = np.float32((bv & 15) - 8)
                        v1 = np.float32((bv >> 4) - 8)
                        dot = np.float32(dot + np.float32(np.float32(v0 * x[b * 32 + j]) + np.float32(v1 * x[b * 32 + j + 16])))
                else:
                    for j in range(32):
                        dot = np.float32(dot + np.float32(np.float32(np.int8(blocks[r, b, 2 + j])) * x[b * 32 + j]))
                s = np.float32(s + np.float32(dot * d))
            exp[r] = s
        assert np.array_equal(got.view(np.uint32), exp.view(np.uint32))


def test_matmul_threading_invariant():
    rng = np.random.default_rng(1)
    rows, cols = 512, 128
    q = G.quantize_q8_0(rng.standard_normal((rows, cols)).astype(np.float32))
    x = rng.standard_normal(cols).astype(np.float32)
    O.set_workers(1)
    a = O.matmul(q, G.GGML_Q8_0, x, rows, cols)
    O.set_workers(7)
    b = O.matmul(q, G.GGML_Q8_0, x, rows, cols)
    O.set_workers(os.cpu_count() or 1)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_unsupported_type_rejected():
    with pytest.raises(ValueError):
        O.dequant(3, np.zeros(20, np.uint8), 32)  # Q4_1: parsed by gguf.go but no Dequant/MatMul exists


FILES = ["tiny_gqa_f16", "tiny_gqa_q8_0", "tiny_gqa_q8_0_requant", "tiny_gqa_q4_0", "tiny_mha_qknorm_q8_0"]


@pytest.mark.parametrize("name", FILES)
def test_forward_logits_vs_reference_torch_model(golden_dir, gold, name):
    gf = G.load_gguf(os.path.join(golden_dir, name + ".gguf"))
    m = O.OracleModel(gf)
    toks = gold["tokens"]
    exp = gold[name + "_logits"]
    m.reset()
    worst = 0.0
    for pos, t in enumerate(toks):
        lg = m.forward(int(t), pos)
        # BASELINE.json tolerance is 1e-3 max-rel; the two fp32 implementations agree far tighter
        err = np.abs(lg - exp[pos]).max() / np.abs(exp[pos]).max()
        worst = max(worst, float(err))
        assert int(np.argmax(lg)) == int(np.argmax(exp[pos]))
    assert worst < 2e-5, worst


@pytest.mark.parametrize("name", FILES)
def test_greedy_stream_vs_reference_torch_model(golden_dir, gold, name):
    gf = G.load_gguf(os.path.join(golden_dir, name + ".gguf"))
    m = O.OracleModel(gf)
    exp = gold[name + "_greedy"]
    got, margins = m.generate_greedy(gold["tokens"][:8], len(exp))
    # identical wherever the top-1/top-2 margin is not at fp32-noise level
    n = len(exp)
    bad = [i for i in range(min(n, len(got))) if got[i] != exp[i]]
    if bad:
        assert margins[bad[0]] < 1e-4, (bad[0], margins[bad[0]])
    else:
        assert len(got) == n


def test_forward_rejects_out_of_range(golden_dir):
    gf = G.load_gguf(os.path.join(golden_dir, "tiny_gqa_q8_0.gguf"))
    m = O.OracleModel(gf)
    with pytest.raises(IndexError):
        m.forward(256, 0)
    with pytest.raises(IndexError):
        m.forward(1, 64)


def test_oracle_samplers_agree_with_the_host_mirror():
    """Two restatements of go/main.go:177-197 / :294-398 -- the C oracle (insertion top-k, qsort top-p) and the host mirror in
    nanollama_b200/engine.py (numpy stable sorts, fp32 loops) -- select the same token for the same random number; the GPU sampler
    (nl_sample) is checked against both in tests/test_gpu_parity.py."""
    from types import SimpleNamespace
    from nanollama_b200.engine import Engine
    rng = np.random.default_rng(12)
    for n, scale in ((1000, 3.0), (4096, 1.0), (257, 8.0)):
        lg = (rng.standard_normal(n) * scale).astype(np.float32)
        lg[7] = lg[3]
        model = SimpleNamespace(state=SimpleNamespace(logits=lg.copy()), config=SimpleNamespace(vocab_size=n))
        host = Engine(model, seed=0)
        for temp, top_k, top_p in [(0.8, 50, 0.9), (1.0, 40, 1.0), (0.7, 1, 1.0), (1.3, 600, 1.0), (0.9, 50, 0.5), (2.0, 50, 0.99), (0.0, 5, 0.9)]:
            for u in (0.0, 0.13, 0.5, 0.77, 0.9991):
                host.rng = SimpleNamespace(random=lambda: u)
                if top_p < 1.0:
                    assert O.sample_top_p(lg, temp, top_p, np.float32(u)) == host.sample_top_p(temp, top_p), (n, temp, top_p, u)
                else:
                    assert O.sample_top_k(lg, temp, top_k, np.float32(u)) == host.sample_top_k(temp, top_k), (n, temp, top_k, u)
    # repetition penalty: once per occurrence, sign-dependent, out-of-range ids ignored
    lg = np.array([2.0, -1.0, 0.0, 4.0], np.float32)
    out = O.rep_penalty(lg, [0, 1, 0, 9, -1, 2], 2.0)
    assert np.array_equal(out, np.array([0.5, -2.0, 0.0, 4.0], np.float32))
    assert np.array_equal(O.rep_penalty(lg, [0, 1], 1.0), lg)
