"""Host-side logic of the tensor-parallel path on the CPU: shard planning rules and the world_size-2 handle
exchange over a gloo process group (the GPU box uses nccl for the same call)."""
import os
import socket
import subprocess
import sys
import textwrap

import pytest

from nanollama_b200 import tiers as T
from nanollama_b200.tp import shard_plan

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_plan_big_all_degrees():
    m = T.tier_meta("big")
    for tp in (1, 2, 4, 8):
        p = shard_plan(m, tp)
        assert p.heads * tp == 64 and p.kv_heads * tp == 16 and p.ffn_rows * tp == 11008 and p.vocab_rows * tp == 96000
        assert p.q_rows % 32 == 0 and p.ffn_rows % 32 == 0          # row-split o/down keeps whole quant blocks
        assert p.heads // p.kv_heads == 4                            # the GQA group stays on one rank
    assert shard_plan(m, 8).allreduces_per_token == 80 and shard_plan(m, 8).allreduce_bytes == 16384
    assert shard_plan(m, 1).allreduces_per_token == 0


def test_shard_plan_rejections():
    with pytest.raises(ValueError, match="n_kv_heads"):
        shard_plan(T.tier_meta("large"), 8)        # 12 KV heads do not split 8 ways
    assert shard_plan(T.tier_meta("large"), 4).kv_heads == 3
    with pytest.raises(ValueError):
        shard_plan(T.tier_meta("mini"), 2)         # 3 KV heads
    with pytest.raises(ValueError):
        shard_plan(T.tier_meta("nano"), 2)         # 9 heads
    with pytest.raises(ValueError, match="supported"):
        shard_plan(T.tier_meta("big"), 3)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def test_handle_exchange_world_size_2_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys
        sys.path.insert(0, {ROOT!r})
        import torch.distributed as dist
        from nanollama_b200.tp import exchange_handles_torch
        dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + os.environ["PORT"], rank=int(os.environ["RANK"]), world_size=2)
        mine = bytes([int(os.environ["RANK"]) + 1]) * 64
        got = exchange_handles_torch(mine)
        assert got == [bytes([1]) * 64, bytes([2]) * 64], got
        dist.barrier(); dist.destroy_process_group()
        print("ok")
    """))
    port = str(_free_port())
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(os.environ, RANK=str(r), PORT=port), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for r in range(2)]
    for p in procs:
        out, err = p.communicate(timeout=120)
        assert p.returncode == 0 and "ok" in out, err[-2000:]
