"""Tensor-parallel plumbing for the big tiers: one process per GPU, torch.distributed only for the control plane.

The data path is inside libnanollama_cuda.so: every rank owns a window of device memory, exports it with CUDA IPC, and after
the handles have been exchanged here the per-layer sums and the logits gather run as one-shot stores into peer memory over
NVLink (csrc/nl_tp.cuh).  No NCCL call sits on the per-token path.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional


@dataclass
class ShardPlan:
    """Column-split q/k/v/gate/up (rows of W), row-split o/down (whole 32-element blocks of the input), vocab-split LM head."""
    tp: int
    heads: int          # query heads per rank
    kv_heads: int
    q_rows: int         # rows of attn_q kept per rank (= columns of attn_output kept)
    kv_rows: int
    ffn_rows: int       # rows of ffn_gate / ffn_up (= columns of ffn_down)
    vocab_rows: int
    allreduces_per_token: int
    allreduce_bytes: int


def shard_plan(meta, tp: int) -> ShardPlan:
    """Validates a tensor-parallel degree for a model (same rules as nl_create) and returns the per-rank shard sizes."""
    if tp not in (1, 2, 4, 8):
        raise ValueError(f"tp_size {tp} (supported: 1, 2, 4, 8)")
    hd = meta.head_dim or (meta.embed_dim // meta.num_heads)
    kv = meta.num_kv_heads or meta.num_heads
    if meta.num_heads % tp or kv % tp:
        raise ValueError(f"tp_size {tp} does not divide n_heads {meta.num_heads} / n_kv_heads {kv} evenly")
    if meta.interm_size % (32 * tp) or (meta.num_heads // tp * hd) % 32:
        raise ValueError(f"tp_size {tp}: shard boundaries would cut a 32-element quant block")
    if meta.vocab_size % (4 * tp):
        raise ValueError(f"tp_size {tp} does not divide vocab_size {meta.vocab_size} into float4-aligned shards")
    return ShardPlan(tp, meta.num_heads // tp, kv // tp, meta.num_heads // tp * hd, kv // tp * hd, meta.interm_size // tp,
                     meta.vocab_size // tp, 2 * meta.num_layers if tp > 1 else 0, 4 * meta.embed_dim)


def exchange_handles_torch(handle: bytes, group=None) -> List[bytes]:
    """All-gather of the 64-byte IPC handles through the already initialised torch.distributed process group
    (nccl on the GPU box, gloo in the CPU tests)."""
    import torch.distributed as dist
    out: List[Optional[bytes]] = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, bytes(handle), group=group)
    return [bytes(h) for h in out]  # index == rank


Exchange = Callable[[bytes], List[bytes]]
