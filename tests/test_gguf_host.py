"""Host logic: GGUF parser / metadata / exact producers (go/gguf.go, scripts/export_gguf.py, quantize_gguf.py)."""
import os
import struct

import numpy as np
import pytest

from nanollama_b200 import gguf as G


def test_parse_reference_written_file(golden_dir):
    gf = G.load_gguf(os.path.join(golden_dir, "tiny_gqa_q4_0.gguf"))
    m = gf.meta
    assert (m.num_layers, m.embed_dim, m.num_heads, m.num_kv_heads, m.head_dim) == (2, 128, 2, 1, 64)
    assert (m.vocab_size, m.seq_len, m.interm_size) == (256, 64, 512)
    assert m.rope_theta == 10000.0 and abs(m.rms_norm_eps - 1e-5) < 1e-12
    assert not m.qk_norm and not m.rope_conjugate and m.bos_id == 1 and m.eos_id == 2
    assert len(gf.tensors) == 3 + 2 * 9
    raw, info = gf.get_tensor("blk.1.ffn_down.weight")
    assert info.type == G.GGML_Q4_0 and info.rows_cols == (128, 512) and raw.size == 128 * 512 // 32 * 18
    for i in gf.tensors.values():
        assert i.offset % 32 == 0
    assert gf.get_tensor("output_norm.weight")[1].type == G.GGML_F32
    with pytest.raises(KeyError):
        gf.get_tensor("nope")
    assert G.load_gguf(os.path.join(golden_dir, "tiny_mha_qknorm_q8_0.gguf")).meta.qk_norm


def test_requant_file_uint32_arrays_accepted(golden_dir):
    # quantize_gguf.py rewrites non-negative int arrays as uint32 (:414-418); toInt accepts any int kind
    gf = G.load_gguf(os.path.join(golden_dir, "tiny_gqa_q8_0_requant.gguf"))
    assert gf.meta.token_types[:4] == [2, 3, 3, 6] and gf.meta.vocab_size == 256
    assert all(t.type in (G.GGML_F32, G.GGML_Q8_0) for t in gf.tensors.values())


def test_bad_files(tmp_path):
    p = tmp_path / "x.gguf"
    p.write_bytes(b"NOPE" + b"\0" * 40)
    with pytest.raises(G.GGUFError, match="bad magic"):
        G.load_gguf(str(p))
    p.write_bytes(struct.pack("<IIQQ", G.GGUF_MAGIC, 4, 0, 0) + b"\0" * 40)
    with pytest.raises(G.GGUFError, match="unsupported GGUF version"):
        G.load_gguf(str(p))
    p.write_bytes(struct.pack("<IIQQ", G.GGUF_MAGIC, 3, 0, 0))
    with pytest.raises(G.GGUFError, match="no tensor data"):
        G.load_gguf(str(p))
    with pytest.raises(G.GGUFError, match="open GGUF"):
        G.load_gguf(str(tmp_path / "missing.gguf"))


def test_quantizers_byte_identical_to_reference_bytes(golden_dir):
    kat = np.load(os.path.join(golden_dir, "dequant_kat.npz"))
    src = kat["src_f32"]
    assert np.array_equal(G.quantize_q4_0(src), kat["q4_0_bytes"])
    assert np.array_equal(G.quantize_q8_0(src), kat["q8_0_bytes"])
    half = src.astype(np.float16).astype(np.float64)
    assert np.array_equal(G.quantize_q8_0(half, flavor="requant"), kat["q8_0_requant_bytes"])


def test_writer_roundtrip_byte_identical_to_reference_writer(golden_dir, tmp_path):
    # re-emit a reference-written file through our writer: must reproduce it byte for byte
    src = os.path.join(golden_dir, "tiny_gqa_q8_0.gguf")
    gf = G.load_gguf(src)
    w = G.GGUFWriter(str(tmp_path / "o.gguf"))
    for k, v in gf.meta.kv.items():
        t = gf.meta.kv_types[k]
        if t == G.T_ARRAY:
            w.add_array(k, v.elem_type, list(v))
        else:
            w.kv.append((k, t, v))
    for name, info in gf.tensors.items():
        raw, _ = gf.get_tensor(name)
        w.add_tensor_raw(name, raw, info.type, tuple(reversed(info.dims)))
    w.write()
    assert open(src, "rb").read() == open(w.path, "rb").read()


# ---------------------------------------------------------------- fast exact producers
def _producer_source(seed, nblocks):
    """tests/golden/make_producer_kat.py:source (kept in sync)."""
    rng = np.random.default_rng(seed)
    t = rng.standard_normal((nblocks, 32)).astype(np.float32)
    t[0] = 0.0
    t[1] = 0.0; t[1, 5] = 3.25
    t[2] = 0.0; t[2, 17] = -7.5
    t[3] *= np.float32(1e-6)
    t[4] *= np.float32(1e3)
    t[5] = np.linspace(-1, 1, 32, dtype=np.float32)
    t[6] = np.float32(6e-8)
    t[7::64] *= np.float32(37.0)
    return t.reshape(-1)


def test_producers_byte_identical_on_4096_blocks(golden_dir):
    """numpy quantizers against the bytes the REFERENCE's tensor_to_q4_0 / tensor_to_q8_0 / quantize_to_q8_0 emitted for the same
    131,072 values (digests + stream heads written by tests/golden/make_producer_kat.py from the reference's own functions)."""
    import hashlib
    kat = np.load(os.path.join(golden_dir, "producer_kat.npz"))
    x = _producer_source(int(kat["seed"]), int(kat["nblocks"]))
    got = {"q4_0": G.quantize_q4_0(x), "q8_0": G.quantize_q8_0(x),
           "q8_0_requant": G.quantize_q8_0(x.astype(np.float16).astype(np.float32), flavor="requant")}
    for k, v in got.items():
        assert v.size == int(kat[k + "_nbytes"]), k
        assert np.array_equal(v[:512], kat[k + "_head"]), k
        assert hashlib.sha256(v.tobytes()).digest() == kat[k + "_sha256"].tobytes(), k


def test_quantize_cli_q8_0_is_the_reference_cli_file(golden_dir, tmp_path):
    """python -m nanollama_b200.quantize in out == python scripts/quantize_gguf.py in out, byte for byte (the golden requant file
    was written by the reference CLI from the same F16 file, tests/golden/make_golden.py)."""
    from nanollama_b200 import quantize as Q
    out = tmp_path / "req.gguf"
    assert Q.main([os.path.join(golden_dir, "tiny_gqa_f16.gguf"), str(out), "-q"]) == 0
    assert out.read_bytes() == open(os.path.join(golden_dir, "tiny_gqa_q8_0_requant.gguf"), "rb").read()


def test_quantize_cli_q4_0_path(golden_dir, tmp_path):
    """--dtype q4_0 (the path export_gguf.py's CLI rejects, :432): every 2-D tensor carries tensor_to_q4_0's block bytes of the F16
    file's values, norms are F32, metadata survives, and the engine (oracle) runs the file."""
    from nanollama_b200 import quantize as Q
    from oracle import oracle as O
    src = G.load_gguf(os.path.join(golden_dir, "tiny_gqa_f16.gguf"))
    out = tmp_path / "q4.gguf"
    assert Q.main([os.path.join(golden_dir, "tiny_gqa_f16.gguf"), str(out), "--dtype", "q4_0", "-q"]) == 0
    gf = G.load_gguf(str(out))
    assert gf.meta.kv["llama.block_count"] == src.meta.kv["llama.block_count"] and gf.meta.vocab_size == src.meta.vocab_size
    for name, info in src.tensors.items():
        raw, i2 = gf.get_tensor(name)
        if info.ndims == 1:
            assert i2.type == G.GGML_F32
        else:
            assert i2.type == G.GGML_Q4_0
            vals = np.asarray(src.get_tensor(name)[0]).view(np.float16).astype(np.float32)
            assert np.array_equal(np.asarray(raw), G.quantize_q4_0(vals)), name
    lg = O.OracleModel(gf).forward(1, 0)
    assert np.isfinite(lg).all() and lg.shape == (gf.meta.vocab_size,)
