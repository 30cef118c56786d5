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
