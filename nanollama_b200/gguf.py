"""GGUF container: parser, config extraction, and fast exact producers.

Host-side mirror of the reference's container layer:
  * ``load_gguf``        -> go/gguf.go:289-414  (LoadGGUF: header, KV, tensor infos, 32-B aligned blob)
  * ``parse_metadata``   -> go/gguf.go:417-558  (parseMetadata: arch prefix, dims, flags, tokenizer)
  * ``GGUFFile.get_tensor`` -> go/gguf.go:561-574 (bounds-checked slice of the blob)
  * ``tensor_bytes``     -> go/gguf.go:238-286
  * ``GGUFWriter``       -> scripts/export_gguf.py:164-311 (same byte layout: dims reversed, every
                            tensor 32-B aligned inside the data section)
  * ``quantize_q4_0`` / ``quantize_q8_0`` -> scripts/export_gguf.py:85-159 and
                            scripts/quantize_gguf.py:183-215, vectorised but byte-identical
                            (the reference producers run at 0.1-0.6 M elem/s; tiers above nano need these).
No compute for the forward path happens here; tensors are handed to libnanollama_cuda.so as raw bytes.
"""
from __future__ import annotations

import mmap
import os
import struct
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Tuple

import numpy as np

GGUF_MAGIC = 0x46554747
GGUF_VERSION = 3
GGUF_ALIGNMENT = 32

# GGUF KV value types (go/gguf.go:27-41)
T_UINT8, T_INT8, T_UINT16, T_INT16, T_UINT32, T_INT32, T_FLOAT32, T_BOOL, T_STRING, T_ARRAY, T_UINT64, T_INT64, T_FLOAT64 = range(13)
_SCALAR_FMT = {T_UINT8: "<B", T_INT8: "<b", T_UINT16: "<H", T_INT16: "<h", T_UINT32: "<I", T_INT32: "<i",
               T_FLOAT32: "<f", T_BOOL: "<B", T_UINT64: "<Q", T_INT64: "<q", T_FLOAT64: "<d"}

# GGML tensor types (go/gguf.go:44-56)
GGML_F32, GGML_F16, GGML_Q4_0, GGML_Q4_1, GGML_Q5_0, GGML_Q5_1, GGML_Q8_0 = 0, 1, 2, 3, 6, 7, 8
GGML_Q4_K, GGML_Q6_K = 12, 14
TYPE_NAMES = {GGML_F32: "F32", GGML_F16: "F16", GGML_Q4_0: "Q4_0", GGML_Q5_0: "Q5_0", GGML_Q8_0: "Q8_0",
              GGML_Q4_K: "Q4_K", GGML_Q6_K: "Q6_K"}
TYPE_IDS = {v.lower(): k for k, v in TYPE_NAMES.items()}


def ggml_block_size(t: int) -> int:
    """bytes per block, go/gguf.go:238-260 (0 = unsupported)."""
    return {GGML_F32: 4, GGML_F16: 2, GGML_Q4_0: 18, GGML_Q4_1: 20, GGML_Q8_0: 34, GGML_Q5_0: 22,
            GGML_Q6_K: 210, GGML_Q4_K: 144}.get(t, 0)


def ggml_block_elements(t: int) -> int:
    """elements per block, go/gguf.go:263-272."""
    if t in (GGML_F32, GGML_F16):
        return 1
    if t in (GGML_Q4_K, GGML_Q6_K):
        return 256
    return 32


@dataclass
class GGUFTensorInfo:
    name: str
    ndims: int
    dims: Tuple[int, ...]  # GGML order: innermost first
    type: int
    offset: int

    @property
    def n_elements(self) -> int:
        n = 1
        for d in self.dims[: self.ndims]:
            n *= d
        return n

    @property
    def rows_cols(self) -> Tuple[int, int]:
        """(rows, cols) in the engine's [out_features, in_features] sense."""
        if self.ndims == 1:
            return 1, self.dims[0]
        return self.n_elements // self.dims[0], self.dims[0]


def tensor_bytes(info: GGUFTensorInfo) -> int:
    """go/gguf.go:275-286."""
    be = ggml_block_elements(info.type)
    return (info.n_elements // be) * ggml_block_size(info.type)


@dataclass
class GGUFMetadata:
    """go/gguf.go:60-90."""
    num_layers: int = 0
    embed_dim: int = 0
    num_heads: int = 0
    num_kv_heads: int = 0
    head_dim: int = 0
    vocab_size: int = 0
    seq_len: int = 0
    interm_size: int = 0
    rms_norm_eps: float = 1e-5
    rope_theta: float = 10000.0
    qk_norm: bool = False
    rope_conjugate: bool = False
    token_list: List[str] = field(default_factory=list)
    token_scores: List[float] = field(default_factory=list)
    token_types: List[int] = field(default_factory=list)
    token_merges: List[str] = field(default_factory=list)
    tokenizer_model: str = "llama"
    bos_id: int = 1
    eos_id: int = 2
    add_space_prefix: bool = True
    kv: Dict[str, Any] = field(default_factory=dict)
    kv_types: Dict[str, Any] = field(default_factory=dict)


class GGUFError(Exception):
    pass


class _Reader:
    def __init__(self, buf):
        self.buf = buf
        self.pos = 0

    def take(self, fmt: str):
        size = struct.calcsize(fmt)
        if self.pos + size > len(self.buf):
            raise GGUFError("unexpected EOF")
        (v,) = struct.unpack_from(fmt, self.buf, self.pos)
        self.pos += size
        return v

    def string(self) -> str:
        n = self.take("<Q")
        if n > 1 << 24:  # go/gguf.go:114 sanity limit
            raise GGUFError(f"string too long: {n}")
        if self.pos + n > len(self.buf):
            raise GGUFError("unexpected EOF")
        s = bytes(self.buf[self.pos:self.pos + n]).decode("utf-8", errors="replace")
        self.pos += n
        return s

    def value(self, vtype: int):
        if vtype == T_STRING:
            return self.string()
        if vtype == T_ARRAY:
            et = self.take("<I")
            n = self.take("<Q")
            if et in _SCALAR_FMT and et != T_BOOL:
                fmt = _SCALAR_FMT[et]
                size = struct.calcsize(fmt)
                if self.pos + size * n > len(self.buf):
                    raise GGUFError("unexpected EOF")
                arr = np.frombuffer(self.buf, dtype=np.dtype(fmt[1]).newbyteorder("<"), count=n, offset=self.pos).tolist()
                self.pos += size * n
                return _TypedList(arr, et)
            return _TypedList([self.value(et) for _ in range(n)], et)
        if vtype == T_BOOL:
            return self.take("<B") != 0
        if vtype in _SCALAR_FMT:
            return self.take(_SCALAR_FMT[vtype])
        raise GGUFError(f"unknown GGUF type: {vtype}")


class _TypedList(list):
    """list that remembers its GGUF element type (so files can be re-written verbatim)."""

    def __init__(self, it, elem_type):
        super().__init__(it)
        self.elem_type = elem_type


def _to_int(v) -> int:
    """go/gguf.go:199-220 toInt: any integer/float kind -> int, else 0."""
    if isinstance(v, bool):
        return 0
    if isinstance(v, (int, float)):
        return int(v)
    return 0


def _to_f32(v) -> float:
    if isinstance(v, bool):
        return 0.0
    if isinstance(v, (int, float)):
        return float(np.float32(v))
    return 0.0


def parse_metadata(kv: Dict[str, Any]) -> GGUFMetadata:
    """go/gguf.go:417-558."""
    m = GGUFMetadata(kv=kv)
    arch = kv.get("general.architecture", "llama")
    if not isinstance(arch, str):
        arch = "llama"
    g = lambda k: kv.get(arch + k)
    if g(".block_count") is not None:
        m.num_layers = _to_int(g(".block_count"))
    if g(".embedding_length") is not None:
        m.embed_dim = _to_int(g(".embedding_length"))
    if g(".attention.head_count") is not None:
        m.num_heads = _to_int(g(".attention.head_count"))
    if g(".attention.head_count_kv") is not None:
        m.num_kv_heads = _to_int(g(".attention.head_count_kv"))
    if g(".feed_forward_length") is not None:
        m.interm_size = _to_int(g(".feed_forward_length"))
    if g(".context_length") is not None:
        m.seq_len = _to_int(g(".context_length"))
    if g(".attention.layer_norm_rms_epsilon") is not None:
        m.rms_norm_eps = _to_f32(g(".attention.layer_norm_rms_epsilon"))
    if g(".rope.freq_base") is not None:
        m.rope_theta = _to_f32(g(".rope.freq_base"))
    if m.num_heads > 0 and m.embed_dim > 0:
        m.head_dim = m.embed_dim // m.num_heads
    if m.num_kv_heads == 0:
        m.num_kv_heads = m.num_heads
    if isinstance(kv.get("nanollama.qk_norm"), bool):
        m.qk_norm = kv["nanollama.qk_norm"]
    if isinstance(kv.get("nanollama.rope_conjugate"), bool):
        m.rope_conjugate = kv["nanollama.rope_conjugate"]
    if isinstance(kv.get("tokenizer.ggml.model"), str):
        m.tokenizer_model = kv["tokenizer.ggml.model"]
    toks = kv.get("tokenizer.ggml.tokens")
    if isinstance(toks, list):
        m.token_list = [t if isinstance(t, str) else "" for t in toks]
        m.vocab_size = len(m.token_list)  # go/gguf.go:497 — the ONLY source of VocabSize
    sc = kv.get("tokenizer.ggml.scores")
    if isinstance(sc, list):
        m.token_scores = [_to_f32(s) for s in sc]
    tt = kv.get("tokenizer.ggml.token_type")
    if isinstance(tt, list):
        m.token_types = [_to_int(t) for t in tt]
    if "tokenizer.ggml.bos_token_id" in kv:
        m.bos_id = _to_int(kv["tokenizer.ggml.bos_token_id"])
    if "tokenizer.ggml.eos_token_id" in kv:
        m.eos_id = _to_int(kv["tokenizer.ggml.eos_token_id"])
    mg = kv.get("tokenizer.ggml.merges")
    if isinstance(mg, list):
        m.token_merges = [x if isinstance(x, str) else "" for x in mg]
    asp = kv.get("tokenizer.ggml.add_space_prefix")
    if isinstance(asp, bool):
        m.add_space_prefix = asp
    elif isinstance(asp, int):
        m.add_space_prefix = asp != 0
    return m


class GGUFFile:
    """go/gguf.go:101-106.  ``tensor_data`` is a read-only memory map of the data blob."""

    def __init__(self, meta: GGUFMetadata, tensors: Dict[str, GGUFTensorInfo], tensor_data, data_offset: int, version: int):
        self.meta = meta
        self.tensors = tensors
        self.tensor_data = tensor_data
        self.data_offset = data_offset
        self.version = version

    def get_tensor(self, name: str) -> Tuple[np.ndarray, GGUFTensorInfo]:
        """go/gguf.go:561-574: raw bytes (uint8 view, no copy) + info; KeyError/GGUFError like the Go errors."""
        info = self.tensors.get(name)
        if info is None:
            raise KeyError(f"tensor not found: {name}")
        size = tensor_bytes(info)
        start, end = info.offset, info.offset + size
        if end > len(self.tensor_data):
            raise GGUFError(f"tensor {name} out of bounds: {start} + {size} > {len(self.tensor_data)}")
        return self.tensor_data[start:end], info

    def find_tensor(self, substr: str) -> Optional[GGUFTensorInfo]:
        for n, i in self.tensors.items():
            if substr in n:
                return i
        return None

    def tensors_in_file_order(self):
        """(name, info) pairs in the order of the file's tensor table (the dict is filled while the table is parsed)."""
        return list(self.tensors.items())


def load_gguf(path: str, verbose: bool = False) -> GGUFFile:
    """go/gguf.go:289-414."""
    try:
        f = open(path, "rb")
    except OSError as e:
        raise GGUFError(f"open GGUF: {e}") from e
    with f:
        size = os.fstat(f.fileno()).st_size
        if size < 24:
            raise GGUFError("read magic: unexpected EOF")
        mm = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ)
    r = _Reader(mm)
    magic = r.take("<I")
    if magic != GGUF_MAGIC:
        raise GGUFError(f"bad magic: 0x{magic:08X} (expected 0x{GGUF_MAGIC:08X})")
    version = r.take("<I")
    if version < 2 or version > 3:
        raise GGUFError(f"unsupported GGUF version: {version}")
    n_tensors = r.take("<Q")
    n_kv = r.take("<Q")
    kv: Dict[str, Any] = {}
    kv_types: Dict[str, Any] = {}
    for i in range(n_kv):
        key = r.string()
        vtype = r.take("<I")
        kv[key] = r.value(vtype)
        kv_types[key] = vtype
    tensors: Dict[str, GGUFTensorInfo] = {}
    for i in range(n_tensors):
        name = r.string()
        nd = r.take("<I")
        if nd > 4:
            raise GGUFError(f"tensor {name}: ndims {nd} > 4")
        dims = tuple(r.take("<Q") for _ in range(nd))
        ttype = r.take("<I")
        off = r.take("<Q")
        tensors[name] = GGUFTensorInfo(name, nd, dims, ttype, off)
    header_end = r.pos
    data_offset = (header_end + GGUF_ALIGNMENT - 1) // GGUF_ALIGNMENT * GGUF_ALIGNMENT
    if size - data_offset <= 0:
        raise GGUFError(f"no tensor data (dataOffset={data_offset}, fileSize={size})")
    data = np.frombuffer(mm, dtype=np.uint8, offset=data_offset)
    meta = parse_metadata(kv)
    meta.kv_types = kv_types
    if verbose:
        print(f"[gguf] version={version} tensors={n_tensors} metadata={n_kv}")
        print(f"[gguf] data offset={data_offset} size={(size - data_offset) / 1024 / 1024:.1f} MB")
        print(f"[gguf] layers={meta.num_layers} dim={meta.embed_dim} heads={meta.num_heads} kv_heads={meta.num_kv_heads} head_dim={meta.head_dim}")
        print(f"[gguf] vocab={meta.vocab_size} seq_len={meta.seq_len} ffn={meta.interm_size} rope_theta={meta.rope_theta:.1f} tokenizer={meta.tokenizer_model}")
    return GGUFFile(meta, tensors, data, data_offset, version)


# ─────────────────────────── exact, vectorised producers ───────────────────────────

def quantize_q4_0(x: np.ndarray) -> np.ndarray:
    """Byte-identical to scripts/export_gguf.py:85-121 (tensor_to_q4_0).

    scale = amax/8 in fp32 (positive; +amax clips to 15), zero block -> 1.0; q = clamp(rint(x/scale)+8, 0, 15);
    fp16 scale; byte j = q[j] | q[j+16]<<4.
    """
    t = np.ascontiguousarray(x, dtype=np.float32).reshape(-1, 32)
    amax = np.abs(t).max(axis=1)
    scales = (amax / np.float32(8.0)).astype(np.float32)
    scales[scales == 0] = np.float32(1.0)
    q = np.clip(np.rint(t / scales[:, None]) + np.float32(8.0), 0, 15).astype(np.uint8)
    out = np.empty((t.shape[0], 18), dtype=np.uint8)
    out[:, 0:2] = scales.astype(np.float16).view(np.uint8).reshape(-1, 2)
    out[:, 2:] = q[:, :16] | (q[:, 16:] << 4)
    return out.reshape(-1)


def quantize_q8_0(x: np.ndarray, flavor: str = "export") -> np.ndarray:
    """Byte-identical Q8_0 producers.

    flavor="export":  scripts/export_gguf.py:124-159 (fp32: scale=amax/127, q=rint(x/scale)).
    flavor="requant": scripts/quantize_gguf.py:183-215 (float64: inv=1/scale, q=round(x*inv); scale packed
                      to fp16 straight from the double).
    """
    if flavor == "export":
        t = np.ascontiguousarray(x, dtype=np.float32).reshape(-1, 32)
        amax = np.abs(t).max(axis=1)
        scales = (amax / np.float32(127.0)).astype(np.float32)
        scales[scales == 0] = np.float32(1.0)
        q = np.clip(np.rint(t / scales[:, None]), -128, 127).astype(np.int8)
        sc16 = scales.astype(np.float16)
    elif flavor == "requant":
        t = np.ascontiguousarray(x, dtype=np.float64).reshape(-1, 32)
        amax = np.abs(t).max(axis=1)
        scales = np.where(amax == 0, 1.0, amax / 127.0)
        inv = 1.0 / scales
        q = np.clip(np.rint(t * inv[:, None]), -128, 127).astype(np.int8)
        sc16 = scales.astype(np.float16)  # double -> half, round-to-nearest-even like struct.pack('<e')
    else:
        raise ValueError(flavor)
    out = np.empty((t.shape[0], 34), dtype=np.uint8)
    out[:, 0:2] = sc16.view(np.uint8).reshape(-1, 2)
    out[:, 2:] = q.view(np.uint8)
    return out.reshape(-1)


def encode_tensor(x: np.ndarray, ggml_type: int) -> np.ndarray:
    """scripts/export_gguf.py:201-211 add_tensor dispatch."""
    if ggml_type == GGML_Q4_0:
        return quantize_q4_0(x)
    if ggml_type == GGML_Q8_0:
        return quantize_q8_0(x)
    if ggml_type == GGML_F16:
        return np.ascontiguousarray(x, dtype=np.float32).astype(np.float16).view(np.uint8).reshape(-1)
    if ggml_type == GGML_F32:
        return np.ascontiguousarray(x, dtype=np.float32).view(np.uint8).reshape(-1)
    raise ValueError(f"cannot encode ggml type {ggml_type}")


class GGUFWriter:
    """Same file layout as scripts/export_gguf.py:164-311; tensors may be given as already-encoded bytes."""

    def __init__(self, path: str):
        self.path = path
        self.version = GGUF_VERSION
        self.kv: List[Tuple[str, int, Any]] = []
        self.tensors: List[Tuple[str, Any, int, Tuple[int, ...]]] = []

    def add_uint32(self, k, v): self.kv.append((k, T_UINT32, v))
    def add_int32(self, k, v): self.kv.append((k, T_INT32, v))
    def add_float32(self, k, v): self.kv.append((k, T_FLOAT32, v))
    def add_bool(self, k, v): self.kv.append((k, T_BOOL, v))
    def add_string(self, k, v): self.kv.append((k, T_STRING, v))
    def add_array(self, k, elem_type, values): self.kv.append((k, T_ARRAY, (elem_type, values)))

    def add_tensor_raw(self, name: str, raw, ggml_type: int, shape: Tuple[int, ...]):
        """raw: bytes-like or a zero-arg callable returning a uint8 array (lazy, for multi-GB files)."""
        self.tensors.append((name, raw, ggml_type, tuple(shape)))

    def add_tensor(self, name: str, x: np.ndarray, ggml_type: int):
        self.tensors.append((name, encode_tensor(x, ggml_type), ggml_type, tuple(x.shape)))

    @staticmethod
    def _nbytes(ggml_type: int, shape) -> int:
        n = 1
        for d in shape:
            n *= d
        return n // ggml_block_elements(ggml_type) * ggml_block_size(ggml_type)

    @staticmethod
    def _wstr(f, s: str):
        b = s.encode("utf-8")
        f.write(struct.pack("<Q", len(b)))
        f.write(b)

    def _wkv(self, f, key, vtype, value):
        self._wstr(f, key)
        f.write(struct.pack("<I", vtype))
        if vtype == T_STRING:
            self._wstr(f, value)
        elif vtype == T_BOOL:
            f.write(struct.pack("<B", 1 if value else 0))
        elif vtype == T_ARRAY:
            et, elems = value
            f.write(struct.pack("<I", et))
            f.write(struct.pack("<Q", len(elems)))
            if et == T_STRING:
                for e in elems:
                    self._wstr(f, e)
            else:
                f.write(np.asarray(elems, dtype=np.dtype(_SCALAR_FMT[et][1]).newbyteorder("<")).tobytes())
        else:
            f.write(struct.pack(_SCALAR_FMT[vtype], value))

    def write(self):
        with open(self.path, "wb") as f:
            f.write(struct.pack("<IIQQ", GGUF_MAGIC, self.version, len(self.tensors), len(self.kv)))
            for key, vtype, value in self.kv:
                self._wkv(f, key, vtype, value)
            off = 0
            offsets = []
            for i, (name, raw, t, shape) in enumerate(self.tensors):
                if i > 0:
                    off = (off + GGUF_ALIGNMENT - 1) // GGUF_ALIGNMENT * GGUF_ALIGNMENT
                offsets.append(off)
                off += self._nbytes(t, shape)
            for i, (name, raw, t, shape) in enumerate(self.tensors):
                self._wstr(f, name)
                f.write(struct.pack("<I", len(shape)))
                for d in reversed(shape):  # GGML dims are innermost-first (export_gguf.py:295)
                    f.write(struct.pack("<Q", d))
                f.write(struct.pack("<IQ", t, offsets[i]))
            self._align(f)
            for name, raw, t, shape in self.tensors:
                self._align(f)
                data = raw() if callable(raw) else raw
                data = np.asarray(data, dtype=np.uint8) if not isinstance(data, (bytes, bytearray, memoryview)) else data
                if len(data) != self._nbytes(t, shape):
                    raise GGUFError(f"{name}: {len(data)} bytes, expected {self._nbytes(t, shape)}")
                f.write(data if isinstance(data, (bytes, bytearray, memoryview)) else data.tobytes())

    @staticmethod
    def _align(f):
        pad = (-f.tell()) % GGUF_ALIGNMENT
        if pad:
            f.write(b"\x00" * pad)
