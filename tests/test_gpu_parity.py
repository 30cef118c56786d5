"""Parity of the CUDA path (through the C ABI) with the CPU oracle and the reference-made golden fixtures.
Run on the B200 box: python -m pytest tests -m gpu.
Tolerances: dequant bit-exact; matmul / logits max-rel <= 1e-3 per BASELINE.json (observed ~1e-6); greedy streams identical."""
import os

import numpy as np
import pytest

from nanollama_b200 import gguf as G
from nanollama_b200 import model as M
from nanollama_b200 import tiers as T
from oracle import oracle as O

pytestmark = pytest.mark.gpu

LOGIT_TOL = 1e-3  # BASELINE.json north_star: max relative error of logits, fp32 accumulate


def maxrel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.fixture(scope="module")
def kat(golden_dir):
    return np.load(os.path.join(golden_dir, "dequant_kat.npz"))


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "golden_logits.npz"))


# ---------------------------------------------------------------- dequant: bit-exact
@pytest.mark.parametrize("key,typ", [("q4_0", G.GGML_Q4_0), ("q8_0", G.GGML_Q8_0), ("q8_0_requant", G.GGML_Q8_0), ("f16", G.GGML_F16)])
def test_dequant_golden_bit_exact(kat, key, typ):
    got = M.dequant(typ, kat[key + "_bytes"], 2048).view(np.uint32)
    assert np.array_equal(got, kat[key + "_expect"])


def test_half2float_all_bit_patterns(kat):
    allh = np.arange(65536, dtype=np.uint16)
    got = M.dequant(G.GGML_F16, allh.view(np.uint8), 65536).view(np.uint32)
    # every pattern, NaN payloads included (the reference's table keeps mant << 13, go/gguf.go:623)
    assert np.array_equal(got, kat["half_all_expect"])


@pytest.mark.parametrize("typ", [G.GGML_Q4_0, G.GGML_Q8_0, G.GGML_F16, G.GGML_F32, G.GGML_Q5_0, G.GGML_Q4_K, G.GGML_Q6_K])
def test_dequant_random_bytes_bit_exact_vs_oracle(typ):
    rng = np.random.default_rng(typ)
    n = 256 * 37
    nbytes = n // G.ggml_block_elements(typ) * G.ggml_block_size(typ)
    raw = rng.integers(0, 256, size=nbytes, dtype=np.uint8)
    # keep fp16 scale fields finite (random bytes hit inf/NaN exponents; NaN payload propagation through a multiply is
    # not something the Go engine defines either)
    bs = G.ggml_block_size(typ)
    blocks = raw.reshape(-1, bs) if typ not in (G.GGML_F16, G.GGML_F32) else None
    if typ in (G.GGML_Q4_0, G.GGML_Q8_0, G.GGML_Q5_0):
        blocks[:, 1] &= 0x7B
    elif typ == G.GGML_Q4_K:
        blocks[:, 1] &= 0x7B; blocks[:, 3] &= 0x7B
    elif typ == G.GGML_Q6_K:
        blocks[:, 209] &= 0x7B
    elif typ == G.GGML_F16:
        raw[1::2] &= 0x7B
    elif typ == G.GGML_F32:
        raw[3::4] &= 0x7E
    got = M.dequant(typ, raw, n).view(np.uint32)
    exp = O.dequant(typ, raw, n).view(np.uint32)
    assert np.array_equal(got, exp)


@pytest.mark.parametrize("key,typ", [("q5_0", G.GGML_Q5_0), ("q4_k", G.GGML_Q4_K), ("q6_k", G.GGML_Q6_K)])
def test_kquant_dequant_and_matmul_vs_gguf_py_kat(golden_dir, key, typ):
    """Q5_0 / Q4_K / Q6_K (go/quant.go:171-484) against the gguf-py known answers (tests/golden/make_kquant_kat.py): dequant bit for
    bit, matmul against a float64 dot with the decoded weights and against the oracle."""
    kq = np.load(os.path.join(golden_dir, "kquant_kat.npz"))
    raw, exp = kq[key + "_bytes"], kq[key + "_expect"]
    assert np.array_equal(M.dequant(typ, raw, exp.size).view(np.uint32), exp)
    cols = 1024
    rows = exp.size // cols
    x = np.random.default_rng(7).standard_normal(cols).astype(np.float32)
    got = M.matmul_dispatch(raw, typ, x, rows, cols)
    ref = exp.view(np.float32).reshape(rows, cols).astype(np.float64) @ x.astype(np.float64)
    assert np.abs(got - ref).max() / np.abs(ref).max() < 1e-5
    assert maxrel(got, O.matmul(raw, typ, x, rows, cols)) < 1e-5


def test_dequant_unsupported_type():
    with pytest.raises(Exception, match="unsupported"):
        M.dequant(3, np.zeros(20, np.uint8), 32)


# ---------------------------------------------------------------- matmulDispatch
SHAPES = [(576, 576), (1536, 576), (576, 1536), (192, 768), (2048, 768), (768, 2048), (1000, 768), (7, 64), (4096, 1536)]


@pytest.mark.parametrize("typ", [G.GGML_Q4_0, G.GGML_Q8_0, G.GGML_F16, G.GGML_F32])
@pytest.mark.parametrize("rows,cols", SHAPES)
def test_matmul_vs_oracle(typ, rows, cols):
    rng = np.random.default_rng(rows * 31 + cols + typ)
    w = (rng.standard_normal((rows, cols)) / np.sqrt(cols)).astype(np.float32)
    raw = G.encode_tensor(w, typ)
    x = rng.standard_normal(cols).astype(np.float32)
    got = M.matmul_dispatch(raw, typ, x, rows, cols)
    exp = O.matmul(raw, typ, x, rows, cols)
    assert maxrel(got, exp) < 1e-5


@pytest.mark.parametrize("typ", [G.GGML_Q5_0, G.GGML_Q4_K, G.GGML_Q6_K])
def test_matmul_kquants_vs_oracle(typ):
    rng = np.random.default_rng(typ)
    rows, cols = 96, 1024
    nbytes = rows * cols // G.ggml_block_elements(typ) * G.ggml_block_size(typ)
    raw = rng.integers(0, 256, size=nbytes, dtype=np.uint8)
    blocks = raw.reshape(-1, G.ggml_block_size(typ))
    for col in {G.GGML_Q5_0: [1], G.GGML_Q4_K: [1, 3], G.GGML_Q6_K: [209]}[typ]:
        blocks[:, col] = (blocks[:, col] & 0x03) | 0x28  # modest fp16 scales
    x = rng.standard_normal(cols).astype(np.float32)
    got = M.matmul_dispatch(raw, typ, x, rows, cols)
    exp = O.matmul(raw, typ, x, rows, cols)
    assert maxrel(got, exp) < 1e-5


@pytest.mark.parametrize("batch", [2, 3, 4, 5, 8])
def test_matmul_small_batch(batch):
    rng = np.random.default_rng(batch)
    rows, cols = 1536, 576
    raw = G.quantize_q4_0((rng.standard_normal((rows, cols)) / 24).astype(np.float32))
    x = rng.standard_normal((batch, cols)).astype(np.float32)
    dm = M.DeviceMatrix(raw, G.GGML_Q4_0, rows, cols)
    got = dm.matmul(x)
    for b in range(batch):
        assert maxrel(got[b], O.matmul(raw, G.GGML_Q4_0, x[b], rows, cols)) < 1e-5
    # linearity (size-independent property): W(a*x0 + x1) == a*W x0 + W x1
    y = dm.matmul(2.0 * x[0] + x[1])
    assert maxrel(y, 2.0 * got[0] + got[1]) < 1e-5


def test_matmul_zero_and_shape_errors():
    raw = G.quantize_q8_0(np.zeros((8, 64), np.float32))
    assert np.array_equal(M.matmul_dispatch(raw, G.GGML_Q8_0, np.ones(64, np.float32), 8, 64), np.zeros(8, np.float32))
    with pytest.raises(Exception):
        M.matmul_dispatch(raw, G.GGML_Q8_0, np.ones(48, np.float32), 8, 48)  # cols % 32 != 0


# ---------------------------------------------------------------- Forward
FILES = ["tiny_gqa_f16", "tiny_gqa_q8_0", "tiny_gqa_q8_0_requant", "tiny_gqa_q4_0", "tiny_mha_qknorm_q8_0"]


@pytest.mark.parametrize("name", FILES)
def test_forward_logits_vs_oracle_and_reference_golden(golden_dir, gold, name):
    gf = G.load_gguf(os.path.join(golden_dir, name + ".gguf"))
    m = M.load_llama_model(gf)
    o = O.OracleModel(gf)
    m.reset(); o.reset()
    for pos, t in enumerate(gold["tokens"]):
        m.forward(int(t), pos)
        exp = o.forward(int(t), pos)
        assert maxrel(m.state.logits, exp) < LOGIT_TOL
        assert maxrel(m.state.logits, exp) < 2e-5           # what fp32 reassociation actually costs
        assert maxrel(m.state.logits, gold[name + "_logits"][pos]) < LOGIT_TOL
        assert int(np.argmax(m.state.logits)) == int(np.argmax(exp))
    m.close()


@pytest.mark.parametrize("conj,qkn", [(True, False), (True, True), (False, True)])
def test_forward_flag_variants_vs_oracle(golden_dir, gold, conj, qkn):
    gf = G.load_gguf(os.path.join(golden_dir, "tiny_gqa_q4_0.gguf"))
    m = M.load_llama_model(gf, rope_conjugate=conj, qk_norm=qkn)
    o = O.OracleModel(gf, rope_conjugate=conj, qk_norm=qkn)
    for pos, t in enumerate(gold["tokens"]):
        m.forward(int(t), pos)
        assert maxrel(m.state.logits, o.forward(int(t), pos)) < 2e-5
    m.close()


@pytest.mark.parametrize("name", FILES)
def test_greedy_stream_identical(golden_dir, gold, name):
    gf = G.load_gguf(os.path.join(golden_dir, name + ".gguf"))
    m = M.load_llama_model(gf)
    o = O.OracleModel(gf)
    prompt = gold["tokens"][:8]
    exp, margins = o.generate_greedy(prompt, 56)
    got = m.generate_greedy(prompt, 56)
    assert len(got) == len(exp) == 56  # 8 + 56 = seq_len 64: runs to the context end
    bad = [i for i in range(56) if got[i] != exp[i]]
    assert not bad or margins[bad[0]] < 1e-5, (bad[:3], margins[bad[0]] if bad else None)
    m.close()


def test_generate_stops_like_reference(golden_dir, gold):
    gf = G.load_gguf(os.path.join(golden_dir, "tiny_gqa_q8_0.gguf"))
    m = M.load_llama_model(gf)
    o = O.OracleModel(gf)
    prompt = gold["tokens"][:8]
    exp, _ = o.generate_greedy(prompt, 200)          # context (64) ends first
    got = m.generate_greedy(prompt, 200)
    assert len(got) == len(exp) and np.array_equal(got, exp)
    eos = int(exp[5])                                 # pretend the 6th generated token is EOS
    exp2, _ = o.generate_greedy(prompt, 200, eos_id=eos)
    got2 = m.generate_greedy(prompt, 200, eos_id=eos)
    assert np.array_equal(got2, exp2) and got2[-1] == eos
    long_prompt = np.resize(gold["tokens"], 100)      # longer than the context: prefill stops at seq_len-1
    exp3, _ = o.generate_greedy(long_prompt, 10)
    got3 = m.generate_greedy(long_prompt, 10)
    assert np.array_equal(got3, exp3) and len(got3) == 1
    m.close()


def test_reset_and_replay(golden_dir, gold):
    gf = G.load_gguf(os.path.join(golden_dir, "tiny_gqa_q4_0.gguf"))
    m = M.load_llama_model(gf)
    toks = gold["tokens"]
    for pos, t in enumerate(toks):
        m.forward(int(t), pos)
    first = m.state.logits.copy()
    m.reset()
    for pos, t in enumerate(toks):
        m.forward(int(t), pos)
    assert np.array_equal(first, m.state.logits)      # deterministic, and Reset really clears state
    m.prefill(toks[:12])                              # short prompts are fed token by token: same kernels, same bits
    for pos in range(12, len(toks)):
        m.forward(int(toks[pos]), pos)
    assert np.array_equal(first, m.state.logits)
    m.reset()
    m.prefill(toks)                                   # 16 tokens: one-pass tensor-core prefill (split-bf16 GEMM, ~8e-6)
    assert maxrel(m.state.logits, first) < 1e-4
    m.close()


def test_forward_errors(golden_dir):
    gf = G.load_gguf(os.path.join(golden_dir, "tiny_gqa_q8_0.gguf"))
    m = M.load_llama_model(gf)
    with pytest.raises(IndexError):
        m.forward(256, 0)
    with pytest.raises(IndexError):
        m.forward(-1, 0)
    with pytest.raises(IndexError):
        m.forward(1, 64)
    m.close()

    class Missing:
        meta = gf.meta
        def get_tensor(self, name):
            if name == "blk.1.ffn_up.weight":
                raise KeyError(f"tensor not found: {name}")
            return gf.get_tensor(name)
    with pytest.raises(RuntimeError, match="load weights:.*blk.1.ffn_up.weight"):
        M.load_llama_model(Missing())


def test_tied_embeddings_fallback(golden_dir, gold):
    gf = G.load_gguf(os.path.join(golden_dir, "tiny_gqa_q8_0.gguf"))

    class Tied:
        meta = gf.meta
        tensors = {k: v for k, v in gf.tensors.items() if k != "output.weight"}
        def get_tensor(self, name):
            if name == "output.weight":
                raise KeyError(f"tensor not found: {name}")
            return gf.get_tensor(name)
    m = M.load_llama_model(Tied())
    o = O.OracleModel(Tied())
    for pos, t in enumerate(gold["tokens"][:6]):
        m.forward(int(t), pos)
        assert maxrel(m.state.logits, o.forward(int(t), pos)) < 2e-5
    m.close()


# ---------------------------------------------------------------- optional attention biases (go/model.go:244-262, :525-527, :591)
@pytest.mark.parametrize("name,path", [("tiny_gqa_q4_0", "decode_tiled_kernel"), ("tiny_gqa_q8_0", "decode_tiled_kernel"), ("tiny_gqa_f16", None)])
def test_bias_tensors_every_decode_path(golden_dir, gold, name, path):
    """A model that carries all four bias tensors through the persistent tiled kernel (Q4_0 / Q8_0), the per-matrix chain (F16 and
    NL_NO_TILED), the one-pass prefill and the device-side greedy loop -- logits against the oracle fed the same tensors."""
    from tests.helpers import WithBias
    gf = WithBias(G.load_gguf(os.path.join(golden_dir, name + ".gguf")), seed=3)
    plain = O.OracleModel(gf.gf)
    m = M.load_llama_model(gf)
    assert m.has_bias and (path is None or m.decode_path == path)
    o = O.OracleModel(gf)
    toks = np.resize(gold["tokens"], 40)
    moved = 0
    for pos, t in enumerate(toks):
        m.forward(int(t), pos)
        exp = o.forward(int(t), pos)
        assert maxrel(m.state.logits, exp) < 2e-5, pos
        moved += int(maxrel(plain.forward(int(t), pos), exp) > 1e-3)
    assert moved >= 30          # the biases really change the logits
    last = exp.copy()
    m.reset()
    m.prefill(toks)             # 40 tokens: the tcgen05 GEMM path with bias epilogues
    assert maxrel(m.state.logits, last) < 1e-4
    exp_s, margins = o.generate_greedy(toks[:8], 40)
    got_s = m.generate_greedy(toks[:8], 40)
    bad = [i for i in range(len(exp_s)) if got_s[i] != exp_s[i]]
    assert len(got_s) == len(exp_s) and (not bad or margins[bad[0]] < 1e-5)
    m.close()


def test_bias_subset_and_chain_path(golden_dir, gold):
    """Only some of the bias tensors present (each is optional on its own), on the per-matrix chain (NL_NO_TILED=1 in a fresh process)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import os, sys, numpy as np\n"
        "sys.path.insert(0, %r)\n"
        "from nanollama_b200 import gguf as G, model as M\n"
        "from oracle import oracle as O\n"
        "from tests.helpers import WithBias\n"
        "gold = np.load(os.path.join(%r, 'golden_logits.npz'))\n"
        "for which in (('attn_q', 'attn_output'), ('attn_k', 'attn_v'), ('attn_q', 'attn_k', 'attn_v', 'attn_output')):\n"
        "    gf = WithBias(G.load_gguf(os.path.join(%r, 'tiny_gqa_q4_0.gguf')), seed=5, which=which)\n"
        "    m = M.load_llama_model(gf); o = O.OracleModel(gf)\n"
        "    assert m.decode_path == os.environ['EXPECT_PATH'], m.decode_path\n"
        "    for pos, t in enumerate(gold['tokens']):\n"
        "        m.forward(int(t), pos); exp = o.forward(int(t), pos)\n"
        "        assert np.abs(m.state.logits - exp).max() / np.abs(exp).max() < 2e-5, (which, pos)\n"
        "    m.close()\n"
        "print('BIAS_OK')\n") % (root, golden_dir, golden_dir)
    for env in ({"NL_NO_TILED": "1", "EXPECT_PATH": "gemv_stream_kernel chain"}, {"EXPECT_PATH": "decode_tiled_kernel"},
                {"NL_TILE_POLL": "0", "EXPECT_PATH": "decode_tiled_kernel"}):
        r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **env), capture_output=True, text=True, timeout=600)
        assert r.returncode == 0 and "BIAS_OK" in r.stdout, (env, r.stdout[-1500:], r.stderr[-1500:])


def test_mixed_tensor_types_gate_up_differ(golden_dir, gold):
    """gate in Q4_0, up in Q8_0, one projection downgraded to F16 (scripts/export_gguf.py:600-602): the per-matrix chain with the
    unfused SwiGLU -- every GEMV must see the normed input (the up projection read a scratch buffer nobody wrote)."""
    from tests.helpers import Retyped
    base = G.load_gguf(os.path.join(golden_dir, "tiny_gqa_q4_0.gguf"))
    gf = Retyped(base, {"blk.0.ffn_up.weight": G.GGML_Q8_0, "blk.1.ffn_up.weight": G.GGML_Q8_0, "blk.1.attn_k.weight": G.GGML_F16})
    m = M.load_llama_model(gf)
    o = O.OracleModel(gf)
    for pos, t in enumerate(gold["tokens"]):
        m.forward(int(t), pos)
        assert maxrel(m.state.logits, o.forward(int(t), pos)) < 2e-5, pos
    m.close()


def test_set_gamma_then_generate_without_a_forward(golden_dir, gold):
    """nl_set_gamma drops the captured graphs; the greedy loop, the sequential prefill and the bench loop launch them directly
    (ADVICE r1): they must have been re-captured."""
    gf = G.load_gguf(os.path.join(golden_dir, "tiny_gqa_q8_0.gguf"))
    m = M.load_llama_model(gf)
    o = O.OracleModel(gf)
    rows = (0.05 * np.random.default_rng(3).standard_normal((2, 128))).astype(np.float32)
    mp = np.full(256, -1, np.int32)
    prompt = gold["tokens"][:8]
    mp[int(prompt[1])] = 0; mp[int(prompt[3])] = 1
    m.set_gamma(rows, mp); o.set_gamma(rows, mp)
    got = m.generate_greedy(prompt, 24)             # no forward in between
    exp, margins = o.generate_greedy(prompt, 24)
    bad = [i for i in range(24) if got[i] != exp[i]]
    assert not bad or margins[bad[0]] < 1e-5
    m.set_gamma(None, None)
    m.prefill(prompt)                               # sequential prefill (< 16 tokens) right after dropping gamma
    plain = O.OracleModel(gf)
    for pos, t in enumerate(prompt):
        e2 = plain.forward(int(t), pos)
    assert maxrel(m.state.logits, e2) < 2e-5
    assert m.bench_decode(int(prompt[0]), 0, 4) > 0
    bad_map = mp.copy(); bad_map[5] = 7             # row index outside the table: rejected, not an out-of-bounds read
    with pytest.raises(Exception):
        m.set_gamma(rows, bad_map)
    m.close()


def test_batch_forward_matches_single(golden_dir, gold):
    gf = G.load_gguf(os.path.join(golden_dir, "tiny_gqa_q4_0.gguf"))
    m1 = M.load_llama_model(gf)
    mb = M.load_llama_model(gf, max_batch=3)
    toks = gold["tokens"]
    seqs = [toks[:10], toks[3:13], toks[5:15]]
    singles = []
    for s in seqs:
        m1.reset()
        for pos, t in enumerate(s):
            m1.forward(int(t), pos)
        singles.append(m1.state.logits.copy())
    for pos in range(10):
        out = mb.forward_batch([int(s[pos]) for s in seqs], [pos] * 3)
    for b in range(3):
        assert maxrel(out[b], singles[b]) < 1e-5
    m1.close(); mb.close()


@pytest.mark.parametrize("name", ["tiny_gqa_q4_0", "tiny_gqa_q8_0", "tiny_gqa_f16"])
def test_batch_forward_through_gemms_matches_single(golden_dir, gold, monkeypatch, name):
    """NL_BATCH_GEMM_MIN: a decode batch as token rows of the tcgen05 GEMMs (tall orientation, split K, q|k|v and gate|up in one launch
    each).  Same logits as the single-sequence path within the GEMM tolerance (bf16 hi/lo planes, three products)."""
    monkeypatch.setenv("NL_BATCH_GEMM_MIN", "2")
    gf = G.load_gguf(os.path.join(golden_dir, name + ".gguf"))
    m1 = M.load_llama_model(gf)
    B = 5
    mb = M.load_llama_model(gf, max_batch=B)
    toks = gold["tokens"]
    seqs = [toks[i:i + 10] for i in range(B)]
    singles = []
    for s in seqs:
        m1.reset()
        for pos, t in enumerate(s):
            m1.forward(int(t), pos)
        singles.append(m1.state.logits.copy())
    for pos in range(10):
        out = mb.forward_batch([int(s[pos]) for s in seqs], [pos] * B)
    for b in range(B):
        assert maxrel(out[b], singles[b]) < 1e-4
    m1.close(); mb.close()


def test_gamma_injection(golden_dir, gold):
    gf = G.load_gguf(os.path.join(golden_dir, "tiny_gqa_q8_0.gguf"))
    m = M.load_llama_model(gf)
    o = O.OracleModel(gf)
    rng = np.random.default_rng(3)
    rows = (0.05 * rng.standard_normal((4, 128))).astype(np.float32)
    mp = np.full(256, -1, np.int32)
    toks = gold["tokens"]
    mp[int(toks[1])] = 0; mp[int(toks[2])] = 3
    m.set_gamma(rows, mp); o.set_gamma(rows, mp)
    for pos, t in enumerate(toks[:5]):
        m.forward(int(t), pos)
        assert maxrel(m.state.logits, o.forward(int(t), pos)) < 2e-5
    m.close()


def test_gamma_npz_written_by_the_reference_end_to_end(golden_dir):
    """gamma_ref.npz (written by scripts/extract_gamma.py's save_gamma_npz, tests/golden/make_gamma_golden.py) -> the go/gamma.go mirror
    -> nl_set_gamma: logits of tokens with and without a gamma row against the oracle fed the same dense rows."""
    from nanollama_b200 import gamma as GM
    gf = G.load_gguf(os.path.join(golden_dir, "tiny_gqa_q8_0.gguf"))
    m = M.load_llama_model(gf)
    o = O.OracleModel(gf)
    g = GM.load_gamma(os.path.join(golden_dir, "gamma_ref.npz"))
    assert GM.attach(m, g)
    o.set_gamma(g.values, g.token_to_row(256))
    plain = O.OracleModel(gf)
    changed = 0
    for pos, t in enumerate([1, 3, 17, 42, 200, 255]):
        m.forward(int(t), pos)
        exp = o.forward(int(t), pos)
        assert maxrel(m.state.logits, exp) < 2e-5
        changed += int(not np.array_equal(exp, plain.forward(int(t), pos)))
    assert changed >= 4   # the gamma rows do change the logits
    assert not GM.attach(m, None)
    m.close()


# ---------------------------------------------------------------- BASELINE.json configs 1 and 2
@pytest.mark.parametrize("tier,typ,n_new", [("nano", G.GGML_Q8_0, 256), ("mini", G.GGML_Q8_0, 256), ("mini", G.GGML_Q4_0, 256)])
def test_tier_greedy_256_identical(tier, typ, n_new):
    """Random-init tier -> 16-token prompt -> 256 greedy tokens: identical stream, logits within 1e-3 (BASELINE north_star)."""
    gf = T.SyntheticGGUF(tier, typ, seed=1, seq_len=512)
    m = M.load_llama_model(gf)
    o = O.OracleModel(gf)
    rng = np.random.default_rng(11)
    prompt = np.concatenate([[1], rng.integers(3, gf.meta.vocab_size, size=15)]).astype(np.int32)
    exp, margins = o.generate_greedy(prompt, n_new)
    got = m.generate_greedy(prompt, n_new)
    assert len(got) == len(exp) == n_new
    bad = [i for i in range(n_new) if got[i] != exp[i]]
    assert not bad or margins[bad[0]] < 1e-4, (bad[:3], float(margins[bad[0]]) if bad else None, float(margins.min()))
    # logits at the end of the run (position 16+255) against the oracle's
    m.reset(); o.reset()
    seq = np.concatenate([prompt, exp[:-1]])
    for pos in (0, 1, 2):
        m.forward(int(seq[pos]), pos)
        assert maxrel(m.state.logits, o.forward(int(seq[pos]), pos)) < LOGIT_TOL
    m.close()


@pytest.mark.parametrize("tier,layers,vocab", [("large", 2, 8192), ("big", 1, 4096), ("goldie", 2, 4096)])
def test_wide_tier_logits_vs_oracle(tier, layers, vocab):
    """Full-width layers of the big tiers (more 16-row groups than SMs in every projection, so every band edge of the persistent
    decode kernel cuts a row group / gate-up pair): logits against the oracle over a few positions, then determinism."""
    gf = T.SyntheticGGUF(tier, G.GGML_Q4_0, seed=3, seq_len=160, vocab=vocab, layers=layers)
    m = M.load_llama_model(gf)
    o = O.OracleModel(gf)
    rng = np.random.default_rng(9)
    toks = np.concatenate([[1], rng.integers(3, vocab, size=5)]).astype(np.int32)
    for pos, t in enumerate(toks):
        m.forward(int(t), pos)
        exp = o.forward(int(t), pos)
        assert maxrel(m.state.logits, exp) < 2e-5
        assert int(np.argmax(m.state.logits)) == int(np.argmax(exp))
    last = m.state.logits.copy()
    # long context: several attention splits whose partials are folded across CTAs (GPU only: determinism + finiteness,
    # plus agreement of the final logits with a replay)
    seq = rng.integers(3, vocab, size=150).astype(np.int32)
    m.reset()
    for pos, t in enumerate(seq):
        m.forward(int(t), pos)
    a = m.state.logits.copy()
    m.reset()
    for pos, t in enumerate(seq):
        m.forward(int(t), pos)
    assert np.array_equal(a, m.state.logits) and np.isfinite(a).all()
    m.reset()
    for pos, t in enumerate(toks):
        m.forward(int(t), pos)
    assert np.array_equal(last, m.state.logits)
    m.close(); o.close()


@pytest.mark.parametrize("tier", ["big", "goldie"])
def test_full_depth_benchmarked_config_vs_oracle(tier):
    """The EXACT model bench.py times (big: 40 layers, 4096 wide, 96000 vocabulary, Q4_0 -- BASELINE configs 4 and 3 at full depth):
    logits against the oracle at three prompt positions, 32 greedy steps fed the oracle's token (argmax identical unless the
    oracle's own top-1 / top-2 margin is a rounding tie), and the device-side greedy loop against the same stream."""
    import bench
    gf = T.SyntheticGGUF(tier, G.GGML_Q4_0, seed=0, seq_len=min(2048, bench.PROMPT_LEN + 256 + 8))
    m = M.load_llama_model(gf)
    assert m.decode_path == "decode_tiled_kernel"
    o = O.OracleModel(gf)
    prompt = bench.bench_prompt(gf.meta.vocab_size)
    worst = 0.0
    for pos, t in enumerate(prompt):
        exp = o.forward(int(t), pos)
        m.forward(int(t), pos)
        if pos in (0, 7, len(prompt) - 1):
            worst = max(worst, maxrel(m.state.logits, exp))
            assert maxrel(m.state.logits, exp) < 2e-5, pos
    stream, pos = [], len(prompt)
    for i in range(32):
        srt = np.sort(exp)
        margin = float((srt[-1] - srt[-2]) / max(abs(srt[-1]), 1e-30))
        tok = int(np.argmax(exp))
        assert int(np.argmax(m.state.logits)) == tok or margin < 1e-4, (i, margin)
        stream.append(tok)
        exp = o.forward(tok, pos)
        m.forward(tok, pos)
        assert maxrel(m.state.logits, exp) < LOGIT_TOL, i
        worst = max(worst, maxrel(m.state.logits, exp))
        pos += 1
    got = m.generate_greedy(prompt, 32)
    bad = [i for i in range(32) if got[i] != stream[i]]
    assert not bad, bad[:3]
    assert worst < 1e-4, worst
    m.close(); o.close()


def test_attention_to_the_end_of_the_context_vs_oracle():
    """Positions 1023 / 1024 / 2046 of the 2048-position context cap (go/model.go:145): up to 16 attention splits per kv head whose
    partials are folded across CTAs, second and third passes per split -- against the oracle; then the 2047-token one-pass prefill
    (the most the engine ever feeds, go/main.go:163) against the oracle's logits at the last position."""
    gf = T.SyntheticGGUF("mini", G.GGML_Q4_0, seed=2, seq_len=2048, vocab=2048, layers=4)
    m = M.load_llama_model(gf)
    o = O.OracleModel(gf)
    rng = np.random.default_rng(4)
    seq = rng.integers(3, 2048, size=2047).astype(np.int32)
    for pos, t in enumerate(seq):
        exp = o.forward(int(t), pos)
        m.forward(int(t), pos)
        if pos in (511, 1023, 1024, 1535, 2046):
            assert maxrel(m.state.logits, exp) < 2e-5, pos
            assert int(np.argmax(m.state.logits)) == int(np.argmax(exp)), pos
    last = exp.copy()
    with pytest.raises(IndexError):
        m.forward(1, 2048)
    m.reset()
    m.prefill(seq)
    assert maxrel(m.state.logits, last) < 1e-4
    m.close(); o.close()


def test_prefill_goldie_2047_tokens_vs_oracle():
    """BASELINE config 3's prefill length on goldie's width (2 layers to keep the CPU side short): last-position logits vs the oracle,
    then decode continues from the prefilled cache."""
    gf = T.SyntheticGGUF("goldie", G.GGML_Q4_0, seed=6, seq_len=2048, vocab=4096, layers=2)
    m = M.load_llama_model(gf)
    o = O.OracleModel(gf)
    rng = np.random.default_rng(8)
    seq = np.concatenate([[1], rng.integers(3, 4096, size=2046)]).astype(np.int32)
    for pos, t in enumerate(seq):
        exp = o.forward(int(t), pos)
    m.prefill(seq)
    assert maxrel(m.state.logits, exp) < 1e-4
    nxt = int(np.argmax(exp))
    m.forward(nxt, 2047)
    assert maxrel(m.state.logits, o.forward(nxt, 2047)) < 1e-4
    m.close(); o.close()


def test_long_context_attention_vs_oracle():
    """mini Q4_0 at positions past 128 / 256: the attention phase runs 2-3 splits per kv head with a second pass per split."""
    gf = T.SyntheticGGUF("mini", G.GGML_Q4_0, seed=2, seq_len=320, vocab=2048)
    m = M.load_llama_model(gf)
    o = O.OracleModel(gf)
    rng = np.random.default_rng(4)
    seq = rng.integers(3, 2048, size=300).astype(np.int32)
    for pos, t in enumerate(seq):
        exp = o.forward(int(t), pos)
        if pos in (0, 63, 64, 127, 128, 129, 191, 192, 255, 256, 257, 299):
            m.forward(int(t), pos)
            assert maxrel(m.state.logits, exp) < 2e-5, pos
        else:
            m.forward(int(t), pos)
    m.close(); o.close()


@pytest.mark.parametrize("env", [{"NL_TILE_POLL": "0"}, {"NL_TILE_POLL": "1", "NL_ATT_CHUNK": "48"}, {"NL_ATT_HPI": "1"}, {"NL_ATT_HPI": "99"}],
                         ids=["barriers", "polled_chunk48", "one_head_per_item", "whole_group_per_item"])
def test_decode_modes_vs_oracle(env):
    """The persistent kernel's switches, each in a fresh process: grid barriers instead of polled activations (the path tensor
    parallelism runs), a finer attention split, and both extremes of the q-heads-per-item distribution -- logits at positions around
    the split boundaries and a greedy stream against the oracle (tests/mode_worker.py)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tests", "mode_worker.py")], env=dict(os.environ, **env), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MODE_PARITY_OK" in r.stdout, (r.stdout[-2000:], r.stderr[-2000:])


# ---------------------------------------------------------------- tensor parallel (needs >= 2 GPUs on the box)
def test_tensor_parallel_2gpu_parity():
    import socket
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(root, "tests", "tp_worker.py")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "TP_PARITY_OK" in r.stdout, (r.stdout[-2000:], r.stderr[-2000:])


# ---------------------------------------------------------------- prefill: tcgen05 GEMM path
@pytest.mark.parametrize("typ", [G.GGML_Q4_0, G.GGML_Q8_0, G.GGML_F16])
@pytest.mark.parametrize("rows,cols,batch", [(128, 64, 16), (1000, 768, 200), (2048, 1536, 384), (520, 1376, 40), (300, 1376, 150), (4096, 4096, 64), (777, 1024, 129)])
def test_matmul_many_rows_tensor_core_path(typ, rows, cols, batch):
    """matmulDispatch with >= 16 activation rows runs as one tcgen05 GEMM (split-bf16, fp32 accumulate in TMEM): both orientations of
    nl_gemm2.cuh (<= 128 rows: weights on the M side, narrow matrices split along K; above: one or two 128-row tiles per CTA), K with
    an odd number of quant blocks (1376 = an 8-way shard of big's down projection), ragged N and T."""
    rng = np.random.default_rng(rows + cols + batch + typ)
    raw = G.encode_tensor((rng.standard_normal((rows, cols)) / np.sqrt(cols)).astype(np.float32), typ)
    x = rng.standard_normal((batch, cols)).astype(np.float32)
    got = M.DeviceMatrix(raw, typ, rows, cols).matmul(x)
    for b in (0, 1, batch // 2, batch - 1):
        assert maxrel(got[b], O.matmul(raw, typ, x[b], rows, cols)) < 1e-4   # observed ~8e-6 (dropped lo*lo term ~2^-17)


@pytest.mark.parametrize("name", ["tiny_gqa_q4_0", "tiny_gqa_q8_0", "tiny_gqa_f16", "tiny_mha_qknorm_q8_0"])
def test_prefill_one_pass_vs_token_by_token(golden_dir, gold, name):
    gf = G.load_gguf(os.path.join(golden_dir, name + ".gguf"))
    m = M.load_llama_model(gf)
    o = O.OracleModel(gf)
    toks = np.resize(gold["tokens"], 40)            # 40 prompt tokens -> the GEMM path (>= 16)
    for pos, t in enumerate(toks):
        exp = o.forward(int(t), pos)
    m.reset()
    m.prefill(toks)
    assert maxrel(m.state.logits, exp) < LOGIT_TOL
    assert maxrel(m.state.logits, exp) < 1e-4
    # the KV cache written by the one-pass prefill must carry on into per-token decode
    nxt = int(np.argmax(exp))
    m.forward(nxt, len(toks))
    assert maxrel(m.state.logits, o.forward(nxt, len(toks))) < 1e-4
    # and a prefill that starts at a non-zero position (chunked prompt) sees the earlier chunk
    m.reset(); o.reset()
    m.prefill(toks[:20]); m.prefill(toks[20:], pos0=20)
    for pos, t in enumerate(toks):
        exp = o.forward(int(t), pos)
    assert maxrel(m.state.logits, exp) < 1e-4
    m.close()


def test_prefill_mini_tier_512_tokens():
    """BASELINE config 2: mini Q4_0, 512-token prefill (oracle checked on a 96-token prefix to keep the CPU side short)."""
    gf = T.SyntheticGGUF("mini", G.GGML_Q4_0, seed=1, seq_len=1024)
    m = M.load_llama_model(gf)
    o = O.OracleModel(gf)
    rng = np.random.default_rng(3)
    toks = np.concatenate([[1], rng.integers(3, gf.meta.vocab_size, size=511)]).astype(np.int32)
    for pos in range(96):
        exp = o.forward(int(toks[pos]), pos)
    m.prefill(toks[:96])
    assert maxrel(m.state.logits, exp) < LOGIT_TOL
    m.reset()
    m.prefill(toks)                                   # full 512: determinism + finite
    a = m.state.logits.copy()
    m.reset(); m.prefill(toks)
    assert np.array_equal(a, m.state.logits) and np.isfinite(a).all()
    m.close()


# ---------------------------------------------------------------- device-side sampling (nl_sample; go/main.go:177-197, :294-398)
class _FixedRng:
    def __init__(self, u):
        self.u = u

    def random(self):
        return self.u


def _sampling_model():
    gf = T.SyntheticGGUF("mini", G.GGML_Q4_0, seed=5, seq_len=96, vocab=4096, layers=2)
    m = M.load_llama_model(gf)
    for pos, t in enumerate([1, 77, 1234, 9]):
        m.forward(t, pos)
    return m


def test_device_sampling_matches_host_samplers():
    """nl_sample on the device-resident logits against the host mirror of sampleTopK / sampleTopP (engine.py, itself a restatement of
    go/main.go:294-398) for the same uniform random number: same token (a disagreement needs the random number within an ulp of a cdf
    step, so at most one is tolerated over the sweep)."""
    from nanollama_b200.engine import Engine
    m = _sampling_model()
    host = Engine(m, seed=0)
    logits = m.state.logits.copy()
    bad, bad_oracle, total = [], [], 0
    for temp, top_k, top_p in [(0.8, 50, 0.9), (1.0, 40, 1.0), (0.7, 5, 1.0), (1.3, 3000, 1.0), (0.9, 50, 0.5), (2.5, 50, 0.99), (0.0, 50, 0.9)]:
        for u in (0.0, 0.013, 0.25, 0.5, 0.77, 0.9991):
            m.state.logits[:] = logits
            host.rng = _FixedRng(u)
            exp = host.sample_top_p(temp, top_p) if top_p < 1.0 else host.sample_top_k(temp, top_k)
            got = m.sample(temp, top_k, top_p, 1.0, [], float(np.float32(u)))
            total += 1
            if got != exp:
                bad.append((temp, top_k, top_p, u, exp, got))
            ora = O.sample_top_p(logits, temp, top_p, np.float32(u)) if top_p < 1.0 else O.sample_top_k(logits, temp, top_k, np.float32(u))
            if ora != exp:   # the C oracle's restatement of the same samplers (tests/test_oracle.py ties the two on the CPU)
                bad_oracle.append((temp, top_k, top_p, u, exp, ora))
    assert len(bad) <= 1, (bad[:5], total)
    assert len(bad_oracle) <= 1, (bad_oracle[:5], total)
    # the logits were not touched (no repetition penalty asked for)
    assert np.array_equal(m.get_logits(), logits)
    m.close()


def test_device_sampling_repetition_penalty_in_place():
    """The penalty is applied to the device logits in place, once per occurrence in the window (go/main.go:177-187), then argmax."""
    m = _sampling_model()
    logits = m.state.logits.copy()
    top = int(np.argmax(logits))
    neg = int(np.argmin(logits))
    recent = [top, neg, top, 4095, top]
    pen = np.float32(1.3)
    exp = logits.copy()
    for t in recent:
        exp[t] = np.float32(exp[t] / pen) if exp[t] > 0 else np.float32(exp[t] * pen)
    got_tok = m.sample(0.0, 50, 0.9, float(pen), recent, 0.0)
    assert np.array_equal(m.get_logits(), exp)
    assert got_tok == int(np.argmax(exp))
    with pytest.raises(Exception):
        m.sample(0.8, 0, 1.0, 1.0, [], 0.5)       # top_k < 1 with top-k sampling
    with pytest.raises(Exception):
        m.sample(0.8, 50, 0.9, 1.0, [], 1.0)      # u outside [0, 1)
    m.close()


@pytest.mark.parametrize("top_p,top_k", [(0.9, 50), (1.0, 40)])
def test_engine_device_sampling_same_stream_as_host_sampling(top_p, top_k):
    """Engine.generate_tokens with the samplers on the device (logits never leave HBM) against the host samplers, same seed: the
    same tokens (the default repetition penalty 1.15 over a 64-token window included)."""
    from nanollama_b200.engine import Engine, GenParams
    m = _sampling_model()
    p = GenParams(max_tokens=40, temperature=0.8, top_p=top_p, top_k=top_k)
    prompt = [1, 5, 99, 1000]
    a = Engine(m, eos_id=-1, seed=123).generate_tokens(prompt, p)
    b = Engine(m, eos_id=-1, seed=123, device_sampling=True).generate_tokens(prompt, p)
    assert len(a) == 40 and a == b, (a, b)
    m.close()


# ---------------------------------------------------------------- continuous batching behind /chat (go/serve.go:56,106-108)
def test_continuous_batcher_matches_single_sequence_generation():
    """Six requests through ContinuousBatcher on a 4-sequence model (joins and leaves at step boundaries, idle rows in between) against
    the same requests run alone through Engine.generate_tokens (GenerateQuiet): the same tokens for the same seed.  The batch rows run
    the small-batch GEMV kernels, the lone sequence the persistent kernel: logits differ by fp32 reassociation only, so a sampled token
    can only differ where the random number falls within ~1e-6 of a cdf step (at most one request may differ)."""
    from nanollama_b200.batcher import ContinuousBatcher
    from nanollama_b200.engine import Engine, GenParams
    gf = T.SyntheticGGUF("mini", G.GGML_Q4_0, seed=5, seq_len=96, vocab=4096, layers=2)
    mb = M.load_llama_model(gf, max_batch=4)
    m1 = M.load_llama_model(gf)
    reqs = [([1, 5, 99, 1000], GenParams(max_tokens=24, temperature=0.8, top_p=0.9, top_k=50), 1),
            ([1, 7], GenParams(max_tokens=40, temperature=1.0, top_p=1.0, top_k=8), 2),
            ([1] + list(range(10, 40)), GenParams(max_tokens=16, temperature=0.0, top_p=0.9, top_k=50), 3),
            ([1, 4000, 17], GenParams(max_tokens=30, temperature=0.7, top_p=0.5, top_k=50), 4),
            ([1, 3], GenParams(max_tokens=200, temperature=0.9, top_p=0.95, top_k=50), 5),       # runs to the end of the context
            ([1, 2222], GenParams(max_tokens=12, temperature=0.8, top_p=1.0, top_k=3), 6)]
    b = ContinuousBatcher(mb, eos_id=-1)
    tickets = []
    for i, (pr, p, s) in enumerate(reqs):
        tickets.append(b.submit(pr, p, seed=s))
        b.step(); b.step()                      # arrivals spread over time
    b.run_until_idle()
    bad = 0
    for (pr, p, s), t in zip(reqs, tickets):
        exp = Engine(m1, eos_id=-1, seed=s).generate_tokens(pr, p)
        got = t.result(0)
        assert len(got) > 0
        bad += int(got != exp)
    assert bad <= 1, bad
    assert b.rows_run / b.steps_run > 1.5
    mb.close(); m1.close()
