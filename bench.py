#!/usr/bin/env python3
"""bench.py — headline benchmark: batch-1 greedy decode tok/s of a random-init reference tier in Q4_0 (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--tier big] [--dtype q4_0] [--impl reference]

One *step* = one pass of the hot path over one batch of synthetic input = `--tokens-per-step` (256) consecutive
batch-1 decode tokens of one sequence, starting after a 16-token prompt (BASELINE.json config 2/4 shape), with device-side
greedy feedback.  `value` is timed with CUDA events with everything resident in HBM; `e2e` is the same loop driven
through the reference-facing call (Forward(token,pos) -> State.Logits on the HOST, host argmax), i.e. token/pos H2D and
the full logits D2H inside the timed region.  `roofline` is for the dominant kernel (the dequant-fused GEMV).
`cpu_baseline` / `--impl reference` time the CPU restatement of the Go engine (oracle/, the Go toolchain is absent) on the
host cores, on a bounded sample (few layers, few tokens) extrapolated to the full tier — a reported baseline only.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PROMPT_LEN = 16


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index), "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference_toks(tier: str, typ: int, n_tokens: int, pos0: int):
    """CPU restatement of the Go engine on the host cores, bounded sample: the tier truncated to 2 and to 6 layers (full
    width, full vocab), n_tokens decode steps each (median per-token time); per-layer and head costs extrapolated linearly to
    the full depth."""
    from nanollama_b200 import tiers as T
    from oracle import oracle as O
    full_layers = T.TIERS[tier][0]
    lo, hi = 2, min(6, full_layers)
    times = {}
    for nl in (lo, hi):
        gf = T.SyntheticGGUF(tier, typ, seed=0, seq_len=PROMPT_LEN + 64, layers=nl)
        o = O.OracleModel(gf)
        o.forward(1, 0)  # warm (page in weights)
        per_tok = []
        for i in range(n_tokens):
            t0 = time.perf_counter()
            o.forward(3 + i, pos0 + i if pos0 + i < PROMPT_LEN + 64 else 1 + i)
            per_tok.append(time.perf_counter() - t0)
        times[nl] = float(np.median(per_tok))
        o.close()
    per_layer = max((times[hi] - times[lo]) / max(hi - lo, 1), 1e-9) if hi > lo else times[lo] / lo
    head = max(times[lo] - lo * per_layer, 0.0)
    t_full = full_layers * per_layer + head
    return 1.0 / t_full, O.get_workers(), (f"{tier} truncated to {lo} and {hi} layers x {n_tokens} tokens (median per-token time), "
                                           f"extrapolated to {full_layers} layers + LM head")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--tier", default="big")
    ap.add_argument("--dtype", default="q4_0", choices=["q4_0", "q8_0", "f16"])
    ap.add_argument("--tokens-per-step", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-tokens", type=int, default=8)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    from nanollama_b200 import gguf as G
    from nanollama_b200 import tiers as T
    typ = G.TYPE_IDS[args.dtype]
    metric = f"decode tok/s ({args.dtype.upper()}, bs=1)"
    config = {"workload": f"{args.tier} {T.TIERS[args.tier]} random-init {args.dtype.upper()} GGUF blocks, batch-1 greedy decode of "
                          f"{args.tokens_per_step} tokens after a {PROMPT_LEN}-token prompt per step",
              "tier": args.tier, "dtype": args.dtype, "tokens_per_step": args.tokens_per_step, "prompt_len": PROMPT_LEN,
              "l2": "weights streamed per token exceed the 126 MB L2 (no flush needed)" if args.tier in ("goldie", "medium", "large", "big")
                    else "weights fit in L2: L2-resident by nature of the tier, stated not flushed"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        vals = []
        for _ in range(max(args.warmup, 0)):
            pass  # the CPU arm needs no warm-up beyond the page-in forward inside cpu_reference_toks
        for _ in range(max(1, min(args.steps, 3))):
            v, cores, sample = cpu_reference_toks(args.tier, typ, args.cpu_tokens, PROMPT_LEN)
            vals.append(v)
        v = float(np.median(vals))
        print(json.dumps({"impl": "reference", "metric": metric, "value": v, "unit": "tok/s", "n_gpus": args.gpus, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": 1000.0 * args.tokens_per_step / v, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": v, "unit": "tok/s", "cores": cores, "kind": "port", "sample": sample},
                          "e2e": {"value": v, "unit": "tok/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return 0

    import torch
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: bench.py has no CPU fallback for the measured arm"}))
        return 1
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from nanollama_b200 import build as B
    if rank == 0:
        B.build()
    if dist:
        dist.barrier()
    from nanollama_b200 import model as M

    tps = args.tokens_per_step
    seq_len = min(2048, PROMPT_LEN + tps + 8)
    gf = T.SyntheticGGUF(args.tier, typ, seed=0, seq_len=seq_len)
    # N > 1: tensor parallel over the N GPUs when the tier shards evenly (big: 2/4/8, large: 2/4), else independent replicas
    tp = 1
    if world > 1:
        from nanollama_b200.tp import shard_plan
        try:
            shard_plan(gf.meta, world)
            tp = world
        except ValueError:
            tp = 1
    t0 = time.time()
    m = M.load_llama_model(gf, device=local_rank, tp_rank=rank if tp > 1 else 0, tp_size=tp)
    load_s = time.time() - t0
    decode_path = m.decode_path
    rng = np.random.default_rng(5)
    prompt = np.concatenate([[1], rng.integers(3, gf.meta.vocab_size, size=PROMPT_LEN - 1)]).astype(np.int32)

    def sync_all():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident arm: K steps, CUDA events inside the library around each step's decode span ----
    m.reset()
    m.prefill(prompt[:-1])
    for _ in range(args.warmup):
        m.bench_decode(int(prompt[-1]), PROMPT_LEN - 1, tps)
    sync_all()
    sampler = ClockSampler(local_rank)
    sampler.start()
    wall0 = time.perf_counter()
    ms_steps = [m.bench_decode(int(prompt[-1]), PROMPT_LEN - 1, tps) for _ in range(args.steps)]
    sync_all()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop()
    ms_total = float(sum(ms_steps))
    if dist:
        t = torch.tensor([ms_total], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    n_seq = 1 if tp > 1 else world   # tensor parallel: one sequence on all GPUs; replicas: every rank decodes its own sequence
    value = n_seq * tps * args.steps / (ms_total / 1000.0)

    # ---- end-to-end arm: Forward(token,pos) -> host logits -> host argmax, per token ----
    def e2e_step():
        m.reset()
        pos = 0
        for t in prompt[:-1]:
            m.forward(int(t), pos); pos += 1
        tok = int(prompt[-1])
        t0 = time.perf_counter()
        for _ in range(tps):
            m.forward(tok, pos); pos += 1
            tok = int(np.argmax(m.state.logits))
        return time.perf_counter() - t0

    e2e_step()
    sync_all()
    e2e_s = sum(e2e_step() for _ in range(max(1, min(args.steps, 3))))
    e2e_steps = max(1, min(args.steps, 3))
    if dist:
        t = torch.tensor([e2e_s], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = n_seq * tps * e2e_steps / e2e_s

    # ---- roofline of the dominant kernel (decode_tiled_kernel: one launch per token, >99 % of the step, profiles/r01_launches_big_decode.md;
    #      models the tiled path does not take fall back to the gemv_stream_kernel chain and report that family instead) ----
    # achieved = algorithmic bytes of one decode step (weights + one embedding row + norm weights + KV read/write at
    # the mid position of the step) / CUDA-event time of that step, i.e. every launch gap, the attention and the argmax kernels are
    # charged to the GEMV too -- a lower bound of the kernel's own bandwidth.  The kernel alone, back to back on the largest
    # matrix of the tier with L2-cold replicas (CUDA events inside nl_matrix_bench), is reported next to it as `kernel_alone`.
    peak, peak_src = load_peaks()
    mid_pos = PROMPT_LEN - 1 + tps // 2
    bytes_tok = T.decode_bytes_per_token(gf.meta, typ, mid_pos)
    weight_only = T.BPE[typ] * T.matmul_params(gf.meta)
    achieved = (bytes_tok / tp) / (ms_per_step / 1000.0 / tps) / 1e9   # per-GPU stream: each rank reads 1/tp of the weights
    alone = None
    try:
        raw, info = gf.get_tensor("output.weight")
        rows, cols = info.rows_cols
        dm = M.DeviceMatrix(raw, typ, rows, cols, device=local_rank)
        nbytes = raw.size
        copies = max(2, int(300e6 // nbytes) + 1)      # rotate over > 2x L2 worth of weights
        ms = dm.bench(batch=1, n_copies=copies, warmup=5, iters=40)
        alone = {"shape": [rows, cols], "bytes": int(nbytes + 4 * cols + 4 * rows), "ms": ms, "GBps": (nbytes + 4 * cols + 4 * rows) / ms / 1e6,
                 "frac": (nbytes + 4 * cols + 4 * rows) / ms / 1e6 / peak}
        dm.close()
    except Exception as e:  # never let the extra measurement kill the bench line
        alone = {"error": str(e)}
    traffic = None
    try:   # measured DRAM bytes per launch of the dominant kernel from the committed ncu capture (one launch = one decoded token)
        tr = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        traffic = tr.get(decode_path, {}).get(f"{args.tier}/{args.dtype}", {}).get("dram_bytes_per_launch")
    except Exception:
        traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_src, "kernel": decode_path,
                "bytes_per_step": int(bytes_tok), "gemv_share_of_bytes": weight_only / bytes_tok, "kernel_alone": alone,
                "launch_us": 1000.0 * ms_per_step / tps,
                "how": "one launch of the persistent kernel = one decoded token: algorithmic bytes per token (SURVEY 8d, mid position of the "
                       "step) / CUDA-event time per token over the timed region (embedding, argmax and memset nodes, <1 % of the step, are "
                       "charged to it); traffic = ncu dram bytes per launch (profiles/r01_traffic.json); kernel_alone = the same kernel "
                       "as a one-phase LM-head GEMV timed back to back on L2-cold replicas"}

    out = {"metric": metric, "value": value, "unit": "tok/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if tp > 1 else "weak", "vs_baseline": None, "dtype": "f32",
           "data": "synthetic", "config": dict(config, parallelism=(f"tp{tp} (column-split q/k/v/gate/up, row-split o/down, vocab-split LM head, one-shot NVLink all-reduce x{2 * gf.meta.num_layers}/token)" if tp > 1 else f"replicas x{world}" if world > 1 else "single GPU"), load_s=round(load_s, 1)),
           "clocks": clocks, "roofline": roofline,
           "e2e": {"value": e2e_value, "unit": "tok/s", "h2d_bytes_per_step": 8 * tps, "d2h_bytes_per_step": 4 * gf.meta.vocab_size * tps},
           "gpu_launches": m.launches_per_token * tps * args.steps, "wall_s": round(wall, 3)}
    m.close()
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, cores, sample = cpu_reference_toks(args.tier, typ, args.cpu_tokens, PROMPT_LEN)
        out["cpu_baseline"] = {"value": v, "unit": "tok/s", "cores": cores, "kind": "port", "sample": sample}
    if rank == 0:
        print(json.dumps(out))
    if dist:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
