"""Model wrappers shared by the parity tests and the torchrun workers: GGUF-like objects (``.meta``, ``.tensors``, ``.get_tensor``)
derived from another one."""
import zlib

import numpy as np

from nanollama_b200 import gguf as G


class WithBias:
    """Adds the four optional attention bias tensors the Go engine loads when present (go/model.go:244-262: attn_q.bias,
    attn_k.bias, attn_v.bias, attn_output.bias; added at model.go:525-527 and :591) -- F32, seeded, large enough to move the logits.
    nanollama's own exporters never write them; third-party GGUFs (Qwen) do."""

    def __init__(self, gf, seed=0, scale=0.25, which=("attn_q", "attn_k", "attn_v", "attn_output")):
        self.gf, self.meta, self.seed, self.scale = gf, gf.meta, seed, scale
        m = gf.meta
        kvd = m.num_kv_heads * m.head_dim
        self.sizes = {"attn_q": m.num_heads * m.head_dim, "attn_k": kvd, "attn_v": kvd, "attn_output": m.embed_dim}
        self.which = tuple(which)
        self.tensors = dict(gf.tensors)
        for i in range(m.num_layers):
            for s in self.which:
                name = f"blk.{i}.{s}.bias"
                self.tensors[name] = G.GGUFTensorInfo(name, 1, (self.sizes[s],), G.GGML_F32, 0)

    def get_tensor(self, name):
        if name.endswith(".bias"):
            info = self.tensors.get(name)
            if info is None:
                raise KeyError(f"tensor not found: {name}")
            rng = np.random.Generator(np.random.PCG64([self.seed, zlib.crc32(name.encode())]))
            return (self.scale * rng.standard_normal(info.n_elements)).astype(np.float32).view(np.uint8), info
        return self.gf.get_tensor(name)


class Retyped:
    """Re-encodes chosen tensors of a model in another GGML type (decode with the host codec, encode again): a file in which the
    exporter downgraded single tensors (scripts/export_gguf.py:600-602), or gate and up of different types."""

    def __init__(self, gf, retype):
        self.gf, self.meta, self.retype = gf, gf.meta, dict(retype)
        self.tensors = dict(gf.tensors)
        for name, t in self.retype.items():
            i = gf.tensors[name]
            self.tensors[name] = G.GGUFTensorInfo(i.name, i.ndims, i.dims, t, 0)

    def get_tensor(self, name):
        raw, info = self.gf.get_tensor(name)
        if name not in self.retype:
            return raw, info
        rows, cols = info.rows_cols
        from oracle import oracle as O   # (test infrastructure decoding a fixture)
        w = O.dequant(info.type, raw, rows * cols).reshape(rows, cols)
        return G.encode_tensor(w, self.retype[name]), self.tensors[name]
