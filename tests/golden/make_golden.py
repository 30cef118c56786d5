#!/usr/bin/env python3
"""Generate the golden fixtures under tests/golden/ FROM THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference; the GPU box has no copy):
    python tests/golden/make_golden.py

What it pins (SURVEY.md §8c — the reference ships no KATs for this path, so these are made from its own code):
  1. dequant_kat.npz   bytes emitted by the reference producers scripts/export_gguf.py:tensor_to_q4_0 /
                       tensor_to_q8_0 and scripts/quantize_gguf.py:quantize_to_q8_0 on seeded tensors (with
                       edge blocks: all-zero, single spike, negative max, tiny values), decoded by the independent
                       gguf-py 0.19 `gguf.quants.dequantize`  -> expected fp32 values (bit patterns).
  2. tiny_*.gguf       complete model files written by the reference's GGUFWriter with the metadata keys of
                       export_gguf.main (:520-537) + a synthetic token list (the Go engine takes VocabSize from it,
                       go/gguf.go:497); tiny_gqa_q8_0_requant.gguf comes from running scripts/quantize_gguf.py.
  3. golden_logits.npz logits of the reference torch model nanollama.llama.Llama (fp32, weights = the GGUF tensors
                       decoded by gguf-py, cos/sin kept in fp32) for a fixed token sequence, plus its greedy streams.
"""
import os
import subprocess
import sys
import tempfile

import numpy as np
import torch

REF = "/root/reference"
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REF, "scripts"))
HERE = os.path.dirname(os.path.abspath(__file__))

import export_gguf as EG  # noqa: E402  (reference producer)
import quantize_gguf as QG  # noqa: E402  (reference re-quantizer)
from gguf import GGUFReader  # noqa: E402  (independent decoder)
from gguf.quants import dequantize  # noqa: E402
from gguf.constants import GGMLQuantizationType as QT  # noqa: E402
from nanollama.llama import Llama, LlamaConfig  # noqa: E402


def kat_tensor(seed, nblocks):
    g = torch.Generator().manual_seed(seed)
    t = torch.randn(nblocks, 32, generator=g)
    t[0] = 0.0                                   # all-zero block -> scale 1.0 path
    t[1] = 0.0; t[1, 5] = 3.25                   # single positive spike (clips to 15 in Q4_0)
    t[2] = 0.0; t[2, 17] = -7.5                  # single negative spike
    t[3] *= 1e-6                                 # tiny values -> subnormal fp16 scale
    t[4] *= 1e3                                  # large values
    t[5] = torch.linspace(-1, 1, 32)             # exact ties in rounding
    t[6] = 6e-8                                  # scale underflows fp16 -> 0
    return t.reshape(-1)


def make_dequant_kat():
    out = {}
    t = kat_tensor(1234, 64)
    q4 = np.frombuffer(EG.tensor_to_q4_0(t), dtype=np.uint8)
    q8 = np.frombuffer(EG.tensor_to_q8_0(t), dtype=np.uint8)
    q8r = np.frombuffer(QG.quantize_to_q8_0([float(v) for v in t.half().float().tolist()]), dtype=np.uint8)
    f16 = np.frombuffer(EG.tensor_to_bytes(t, torch.float16), dtype=np.uint8)
    out["src_f32"] = t.numpy()
    out["q4_0_bytes"] = q4
    out["q4_0_expect"] = dequantize(q4, QT.Q4_0).astype(np.float32).view(np.uint32)
    out["q8_0_bytes"] = q8
    out["q8_0_expect"] = dequantize(q8, QT.Q8_0).astype(np.float32).view(np.uint32)
    out["q8_0_requant_bytes"] = q8r
    out["q8_0_requant_expect"] = dequantize(q8r, QT.Q8_0).astype(np.float32).view(np.uint32)
    out["f16_bytes"] = f16
    out["f16_expect"] = f16.view(np.float16).astype(np.float32).view(np.uint32)
    # every fp16 bit pattern through numpy's IEEE conversion (go/gguf.go:601-636 LUT check)
    allh = np.arange(65536, dtype=np.uint16)
    out["half_all_expect"] = allh.view(np.float16).astype(np.float32).view(np.uint32)
    np.savez_compressed(os.path.join(HERE, "dequant_kat.npz"), **out)
    print("dequant_kat.npz:", {k: v.shape for k, v in out.items()})


def build_model(cfg, seed):
    torch.manual_seed(seed)
    m = Llama(cfg)
    m.init_weights()
    s = (3 ** 0.5) * (cfg.n_embd ** -0.5)
    with torch.no_grad():
        # init_weights zeroes c_proj/down_proj and uses std 1e-3 for output -> degenerate; re-init
        for layer in m.layers:
            torch.nn.init.uniform_(layer.attn.c_proj.weight, -s, s)
            torch.nn.init.uniform_(layer.ffn.down_proj.weight, -s, s)
            layer.attn_norm.weight.copy_(1 + 0.1 * torch.randn(cfg.n_embd))
            layer.ffn_norm.weight.copy_(1 + 0.1 * torch.randn(cfg.n_embd))
        torch.nn.init.uniform_(m.output.weight, -s, s)
        m.norm.weight.copy_(1 + 0.1 * torch.randn(cfg.n_embd))
    return m.float()


def write_gguf(model, cfg, path, ggml_type):
    """Reference GGUFWriter + the KV set of export_gguf.main (:520-537), plus a synthetic token list."""
    state = {k: v.detach() for k, v in model.state_dict().items()}
    head_dim = cfg.n_embd // cfg.n_head
    w = EG.GGUFWriter(path)
    w.add_string("general.architecture", "llama")
    w.add_string("general.name", "nanollama-" + os.path.splitext(os.path.basename(path))[0])
    w.add_uint32("llama.block_count", cfg.n_layer)
    w.add_uint32("llama.embedding_length", cfg.n_embd)
    w.add_uint32("llama.attention.head_count", cfg.n_head)
    w.add_uint32("llama.attention.head_count_kv", cfg.n_kv_head)
    w.add_uint32("llama.attention.key_length", head_dim)
    w.add_uint32("llama.attention.value_length", head_dim)
    w.add_uint32("llama.feed_forward_length", EG.compute_intermediate_size(cfg.n_embd))
    w.add_uint32("llama.context_length", cfg.sequence_len)
    w.add_float32("llama.attention.layer_norm_rms_epsilon", cfg.norm_eps)
    w.add_float32("llama.rope.freq_base", cfg.rope_theta)
    w.add_uint32("llama.vocab_size", cfg.vocab_size)
    w.add_bool("nanollama.qk_norm", cfg.use_qk_norm)
    w.add_bool("nanollama.rope_conjugate", False)
    w.add_string("tokenizer.ggml.model", "llama")
    toks = ["<unk>", "<s>", "</s>"] + [f"<0x{i:02X}>" for i in range(cfg.vocab_size - 3)]
    w.add_string_array("tokenizer.ggml.tokens", toks)
    w.add_float32_array("tokenizer.ggml.scores", [0.0] * cfg.vocab_size)
    w.add_int32_array("tokenizer.ggml.token_type", [2, 3, 3] + [6] * (cfg.vocab_size - 3))
    w.add_uint32("tokenizer.ggml.bos_token_id", 1)
    w.add_uint32("tokenizer.ggml.eos_token_id", 2)
    for name in sorted(state.keys()):
        t = state[name]
        w.add_tensor(EG.map_name(name), t, EG.GGML_TYPE_F32 if t.dim() == 1 else ggml_type)
    w.write()


INV = {v: k for k, v in EG.WEIGHT_MAP.items()}
LINV = {v: k for k, v in EG.LAYER_WEIGHT_MAP.items()}


def torch_model_from_gguf(path, cfg):
    """Rebuild the torch reference model from a GGUF using gguf-py's decoder (independent of our code)."""
    r = GGUFReader(path)
    sd = {}
    for t in r.tensors:
        arr = dequantize(t.data, t.tensor_type).astype(np.float32)
        shape = tuple(int(d) for d in reversed(t.shape.tolist()))
        name = t.name
        if name in INV:
            key = INV[name]
        else:
            _, idx, rest = name.split(".", 2)
            key = f"layers.{idx}.{LINV[rest]}"
        sd[key] = torch.from_numpy(arr.reshape(shape).copy())
    m = Llama(cfg).float()
    m.load_state_dict(sd, strict=True)
    # keep RoPE tables in fp32 (llama.py:99-100 rounds them to bf16, which the Go engine does not)
    hd = cfg.n_embd // cfg.n_head
    inv_freq = 1.0 / (cfg.rope_theta ** (torch.arange(0, hd, 2, dtype=torch.float32) / hd))
    fr = torch.outer(torch.arange(m.rotary_seq_len, dtype=torch.float32), inv_freq)
    m.cos = fr.cos()[None, :, None, :]
    m.sin = fr.sin()[None, :, None, :]
    return m.eval()


def main():
    make_dequant_kat()
    golden = {}
    gqa = LlamaConfig(sequence_len=64, vocab_size=256, n_layer=2, n_head=2, n_kv_head=1, n_embd=128)
    mha = LlamaConfig(sequence_len=64, vocab_size=256, n_layer=2, n_head=2, n_kv_head=2, n_embd=128, use_qk_norm=True)
    rng = np.random.default_rng(7)
    seq = np.concatenate([[1], rng.integers(3, 256, size=15)]).astype(np.int32)
    golden["tokens"] = seq
    jobs = []
    m_gqa = build_model(gqa, 0)
    for nm, t in (("f16", EG.GGML_TYPE_F16), ("q8_0", EG.GGML_TYPE_Q8_0), ("q4_0", EG.GGML_TYPE_Q4_0)):
        p = os.path.join(HERE, f"tiny_gqa_{nm}.gguf")
        write_gguf(m_gqa, gqa, p, t)
        jobs.append((f"tiny_gqa_{nm}", p, gqa))
    # the reference's own re-quantizer, run as the CLI it is
    p_rq = os.path.join(HERE, "tiny_gqa_q8_0_requant.gguf")
    subprocess.run([sys.executable, os.path.join(REF, "scripts", "quantize_gguf.py"),
                    os.path.join(HERE, "tiny_gqa_f16.gguf"), p_rq], check=True, stdout=subprocess.DEVNULL)
    jobs.append(("tiny_gqa_q8_0_requant", p_rq, gqa))
    m_mha = build_model(mha, 1)
    p = os.path.join(HERE, "tiny_mha_qknorm_q8_0.gguf")
    write_gguf(m_mha, mha, p, EG.GGML_TYPE_Q8_0)
    jobs.append(("tiny_mha_qknorm_q8_0", p, mha))

    for name, path, cfg in jobs:
        tm = torch_model_from_gguf(path, cfg)
        with torch.no_grad():
            logits = tm(torch.from_numpy(seq.astype(np.int64))[None])[0].numpy().astype(np.float32)
            stream = list(tm.generate([int(t) for t in seq[:8]], max_tokens=40, temperature=0.0))
        golden[name + "_logits"] = logits
        golden[name + "_greedy"] = np.asarray(stream, dtype=np.int32)
        srt = np.sort(logits, axis=1)
        print(f"{name}: logits {logits.shape}, min top1-top2 margin {float((srt[:, -1] - srt[:, -2]).min()):.4f}, "
              f"greedy head {stream[:8]}")
    np.savez_compressed(os.path.join(HERE, "golden_logits.npz"), **golden)


if __name__ == "__main__":
    main()
