"""Gamma essence loader: the host mirror of go/gamma.go (LoadGamma :37-276, ApplyToEmbedding :279-299) and of the NPZ layouts it
accepts -- the simple one (``indices.npy`` + ``values.npy``) and the sparse COO one written by scripts/extract_gamma.py:143-200
(``tok_embeddings.weight.{indices_0,indices_1,values,shape}``).  The injection itself runs on the device: ``attach`` hands the dense
rows and the token -> row map to ``nl_set_gamma`` (the embedding kernel adds row ``map[token]`` to the dequantised embedding,
go/model.go:502-505)."""
from __future__ import annotations

import zipfile
from dataclasses import dataclass, field
from typing import Dict, Optional

import numpy as np


class GammaError(Exception):
    pass


@dataclass
class GammaEssence:
    """go/gamma.go:22-35.  ``values`` is always fp32 here ([num_tokens, embed_dim]); an f16 file is widened exactly (half2float),
    which is what ApplyToEmbedding does element by element."""
    vocab_size: int
    embed_dim: int
    num_tokens: int
    indices: np.ndarray            # int32 [num_tokens]: token ids that carry a gamma row
    values: np.ndarray             # float32 [num_tokens, embed_dim]
    is_f16: bool
    index_map: Dict[int, int] = field(default_factory=dict)   # token id -> row

    def apply_to_embedding(self, embed: np.ndarray, token: int) -> None:
        """go/gamma.go:279-299: embed += gamma[token] in place (nothing for a token without a row)."""
        row = self.index_map.get(int(token))
        if row is not None:
            embed += self.values[row]

    def token_to_row(self, vocab_size: int) -> np.ndarray:
        """Dense int32 map for nl_set_gamma: row index per token id, -1 = no gamma.  Rows of tokens beyond the model's vocabulary
        are unreachable (the engine never looks them up) and are left out."""
        mp = np.full(vocab_size, -1, dtype=np.int32)
        for tok, row in self.index_map.items():
            if 0 <= tok < vocab_size:
                mp[tok] = row
        return mp


def _member(z: zipfile.ZipFile, name: str) -> Optional[np.ndarray]:
    if name not in z.namelist():
        return None
    with z.open(name) as f:
        return np.lib.format.read_array(f, allow_pickle=False)


def load_gamma(path: str, verbose: bool = False) -> GammaEssence:
    try:
        z = zipfile.ZipFile(path)
    except (OSError, zipfile.BadZipFile) as e:
        raise GammaError(f"open gamma npz: {e}") from e
    with z:
        names = z.namelist()
        if "indices.npy" in names:                                   # format 1 (go/gamma.go:71-76)
            ind_name, val_name = "indices.npy", "values.npy"
        elif "tok_embeddings.weight.indices_0.npy" in names:         # format 2 (:77-81)
            ind_name, val_name = "tok_embeddings.weight.indices_0.npy", "tok_embeddings.weight.values.npy"
        else:
            raise GammaError("gamma npz: no indices found (expected indices.npy or tok_embeddings.weight.indices_0.npy)")
        indices = _member(z, ind_name)
        values = _member(z, val_name)
        if indices is None or values is None:
            raise GammaError("gamma npz missing indices or values")
        indices = np.ascontiguousarray(indices).astype(np.int32).reshape(-1)
        is_f16 = values.dtype == np.float16
        if values.dtype not in (np.float16, np.float32, np.float64):
            raise GammaError(f"read {val_name}: unsupported dtype {values.dtype}")
        values = values.astype(np.float32)                           # f16 -> f32 is exact (half2float)
        indices1 = _member(z, "tok_embeddings.weight.indices_1.npy") if ind_name.startswith("tok_embeddings") else None
        if indices1 is not None:
            # sparse COO [row, col] -> dense [unique tokens, embed_dim], rows in ascending token order (:160-218)
            shape = _member(z, "tok_embeddings.weight.shape.npy")
            if shape is None or shape.size < 2 or int(shape.reshape(-1)[1]) == 0:
                raise GammaError("gamma: tok_embeddings.weight.shape.npy missing or invalid")
            embed_dim = int(shape.reshape(-1)[1])
            indices1 = indices1.astype(np.int64).reshape(-1)
            flat = values.reshape(-1)
            if not (indices.size == indices1.size == flat.size):
                raise GammaError(f"gamma: {indices.size} row indices, {indices1.size} column indices, {flat.size} values")
            tokens = np.unique(indices)                               # sorted ascending: the reference sorts for determinism
            row_of = np.searchsorted(tokens, indices)
            dense = np.zeros((tokens.size, embed_dim), dtype=np.float32)
            if indices1.size and (indices1.min() < 0 or indices1.max() >= embed_dim):
                raise GammaError("gamma: column index out of range")
            dense[row_of, indices1] = flat                            # later duplicates overwrite earlier ones, like the Go loop
            indices, values = tokens.astype(np.int32), dense
        if values.ndim != 2:
            raise GammaError(f"read {val_name}: expected a 2-D array, got shape {values.shape}")
        num_tokens, embed_dim = values.shape
        if indices.size != num_tokens:
            raise GammaError(f"indices len {indices.size} != values rows {num_tokens}")
        index_map = {int(t): i for i, t in enumerate(indices)}       # a repeated token id keeps its LAST row (:226-232)
        vocab = int(indices.max()) + 1 if num_tokens else 0
        if verbose:
            print(f"[gamma] loaded {'f16' if is_f16 else 'f32'}: {num_tokens}/{vocab} tokens, embed_dim={embed_dim} "
                  f"({values.size * (2 if is_f16 else 4) / 1024 / 1024:.1f} MB RAM)")
        return GammaEssence(vocab, embed_dim, num_tokens, indices, np.ascontiguousarray(values), is_f16, index_map)


def attach(model, gamma: Optional[GammaEssence]) -> bool:
    """go/main.go:69-83: keep the gamma only when its embed_dim matches the model's (a mismatch is a warning, not an error); pushes
    the rows to the device.  Returns whether the gamma is active."""
    if gamma is None:
        model.set_gamma(None, None)
        return False
    if gamma.embed_dim != model.config.embed_dim:
        print(f"warning: gamma embed_dim {gamma.embed_dim} != model dim {model.config.embed_dim}, skipping")
        return False
    if gamma.num_tokens == 0:
        model.set_gamma(None, None)
        return False
    model.set_gamma(gamma.values, gamma.token_to_row(model.config.vocab_size))
    return True
