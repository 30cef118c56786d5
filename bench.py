#!/usr/bin/env python3
"""bench.py — headline benchmark: batch-1 greedy decode tok/s of a random-init reference tier in Q4_0 (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--tier big] [--dtype q4_0] [--impl reference] [--mode decode|prefill]

decode (default).  One *step* = one pass of the hot path over one batch of synthetic input = `--tokens-per-step` (256)
consecutive batch-1 decode tokens of one sequence, starting after a 16-token prompt (BASELINE.json config 2/4 shape), with
device-side greedy feedback.  `value` is timed with CUDA events with everything resident in HBM; `e2e` is the same loop driven
through the reference-facing call (Forward(token,pos) -> State.Logits on the HOST, host argmax), i.e. token/pos H2D and the full
logits D2H inside the timed region.  `roofline` is for the dominant kernel (the persistent dequant-fused decode kernel).
`parity_check` (printed at every N): the loaded model against the CPU oracle -- the same tier truncated to 2 layers (logits at
four positions + 16 greedy tokens; tensor parallel: through the same sharded path), the full-depth greedy stream identical on
every rank, and at N = 1 the full-depth logits at the positions the CPU baseline leg computes anyway.

`cpu_baseline` / `--impl reference`: the CPU restatement of the Go engine (oracle/, kind "port": the image has no Go toolchain)
on the host cores, on the FULL-depth model: a step = `--cpu-tokens` real Forward calls at consecutive positions after the
prompt length; nothing is extrapolated, `ms_per_step` is what was timed.

prefill (`--mode prefill`, secondary line): one step = one nl_prefill of `--prefill-tokens` (2047) tokens of goldie Q4_0
(BASELINE config 3) on the tcgen05 GEMM path; roofline against the measured bf16 tensor peak with SURVEY 8d's useful FLOPs.
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


def bench_prompt(vocab: int) -> np.ndarray:
    """The 16-token synthetic prompt of every decode step: BOS + 15 seeded ids in [3, vocab)."""
    rng = np.random.default_rng(5)
    return np.concatenate([[1], rng.integers(3, vocab, size=PROMPT_LEN - 1)]).astype(np.int32)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return {"hbm": float(d["hbm_gbs"]), "tensor": float(d["bf16_tflops"]), "tensor_sustained": float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                    "src": "measured (MEASURED_PEAKS.json)"}
        except Exception:
            pass
    return {"hbm": 6650.0, "tensor": 1590.0, "tensor_sustained": 1400.0, "src": "fallback (B200_PROFILING.md 6.65 TB/s / 1.59 PFLOP/s)"}


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


def decode_config(args, T):
    """The `config` object of a decode line -- built by ONE function so that both arms print the identical dict."""
    return {"workload": f"{args.tier} {T.TIERS[args.tier]} random-init {args.dtype.upper()} GGUF blocks, batch-1 greedy decode of "
                        f"{args.tokens_per_step} tokens after a {PROMPT_LEN}-token prompt per step",
            "tier": args.tier, "dtype": args.dtype, "tokens_per_step": args.tokens_per_step, "prompt_len": PROMPT_LEN,
            "l2": "weights streamed per token exceed the 126 MB L2 (no flush needed)" if args.tier in ("goldie", "medium", "large", "big")
                  else "weights fit in L2: L2-resident by nature of the tier, stated not flushed"}


class CpuEngine:
    """The CPU port of the Go engine on the FULL-depth model, timed in steps of `n_tokens` real Forward calls at consecutive
    positions from PROMPT_LEN on (KV cache rows below are whatever earlier calls left: attention cost depends on the position only)."""

    def __init__(self, gf):
        from oracle import oracle as O
        self.O = O
        self.gf = gf
        self.o = O.OracleModel(gf)
        self.cores = O.get_workers()
        self.o.forward(1, 0)   # page the weights in
        self.cap = min(gf.meta.seq_len, 2048)

    def step(self, tokens, pos0):
        """-> (seconds, [logits copies])"""
        out = []
        t0 = time.perf_counter()
        for i, t in enumerate(tokens):
            out.append(self.o.forward(int(t), pos0 + i).copy())
        return time.perf_counter() - t0, out

    def close(self):
        self.o.close()


def config1_cpu(O_mod, T, G):
    """BASELINE config 1 (the reference's own CPU-runnable case): nano Q8_0, 16-token prompt, 256 greedy tokens on the CPU port;
    timer started after the prefill like go/main.go:171,222-227."""
    gf = T.SyntheticGGUF("nano", G.GGML_Q8_0, seed=0, seq_len=PROMPT_LEN + 256 + 8)
    o = O_mod.OracleModel(gf)
    prompt = bench_prompt(gf.meta.vocab_size)
    for pos, t in enumerate(prompt):
        lg = o.forward(int(t), pos)
    t0 = time.perf_counter()
    pos = len(prompt)
    for _ in range(256):
        tok = int(np.argmax(lg))
        lg = o.forward(tok, pos)
        pos += 1
    dt = time.perf_counter() - t0
    o.close()
    return {"workload": "nano 89M Q8_0, 16-token prompt, 256 greedy tokens, CPU port of the Go engine (timer after prefill, go/main.go:171)",
            "tok_s": 256 / dt, "seconds": dt, "cores": O_mod.get_workers()}


def sample_tokens(vocab, n):
    return (3 + (np.arange(n, dtype=np.int64) * 7919) % (vocab - 3)).astype(np.int32)


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (kind "port") on the host cores, full-depth model,
    `--steps` timed steps of `--cpu-tokens` Forward calls each after `--warmup` untimed ones."""
    from nanollama_b200 import gguf as G
    from nanollama_b200 import tiers as T
    typ = G.TYPE_IDS[args.dtype]
    gf = T.SyntheticGGUF(args.tier, typ, seed=0, seq_len=min(2048, PROMPT_LEN + args.tokens_per_step + 8))
    t0 = time.time()
    eng = CpuEngine(gf)
    load_s = time.time() - t0
    n = args.cpu_tokens
    toks = sample_tokens(gf.meta.vocab_size, n)
    span = max(eng.cap - PROMPT_LEN - n, 1)
    for w in range(max(args.warmup, 0)):
        eng.step(toks, PROMPT_LEN + (w * n) % span)
    secs = []
    for k in range(args.steps):
        s, _ = eng.step(toks, PROMPT_LEN + ((args.warmup + k) * n) % span)
        secs.append(s)
    eng.close()
    total = float(sum(secs))
    v = n * args.steps / total
    line = {"impl": "reference", "metric": f"decode tok/s ({args.dtype.upper()}, bs=1)", "value": v, "unit": "tok/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 * total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": decode_config(args, T),
            "cpu_baseline": {"value": v, "unit": "tok/s", "cores": eng.cores, "kind": "port",
                             "sample": f"full-depth {args.tier} ({gf.meta.num_layers} layers), each step = {n} real Forward calls at consecutive positions "
                                       f">= {PROMPT_LEN}; {args.steps} timed steps after {args.warmup} warm-up steps; nothing extrapolated",
                             "step_tok_s_min_max": [n / max(secs), n / min(secs)], "load_s": round(load_s, 1)},
            "e2e": {"value": v, "unit": "tok/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "sample_tokens_per_step": n}
    if not args.no_config1:
        from oracle import oracle as O
        line["config1_cpu"] = config1_cpu(O, T, G)
    print(json.dumps(line))
    return 0


def run_prefill(args, torch, dist, rank, world, local_rank):
    """Secondary line: one step = one nl_prefill of --prefill-tokens tokens (goldie Q4_0 by default, BASELINE config 3)."""
    from nanollama_b200 import gguf as G
    from nanollama_b200 import model as M
    from nanollama_b200 import tiers as T
    tier = args.tier if args.tier_given else "goldie"
    typ = G.TYPE_IDS[args.dtype]
    Tn = args.prefill_tokens
    gf = T.SyntheticGGUF(tier, typ, seed=0, seq_len=2048)
    tp = 1
    if world > 1:
        from nanollama_b200.tp import shard_plan
        try:
            shard_plan(gf.meta, world); tp = world
        except ValueError:
            tp = 1
    m = M.load_llama_model(gf, device=local_rank, tp_rank=rank if tp > 1 else 0, tp_size=tp)
    rng = np.random.default_rng(0)
    toks = np.concatenate([[1], rng.integers(3, gf.meta.vocab_size, size=Tn - 1)]).astype(np.int32)

    def sync_all():
        torch.cuda.synchronize()
        if dist:
            dist.barrier(); torch.cuda.synchronize()

    # parity before timing, on every rank: the one-pass prefill of a 96-token prefix (tcgen05 GEMMs + tensor-core attention; under tensor
    # parallelism the row exchanges) against the same tokens fed one Forward at a time like go/main.go:160-166 (the persistent decode
    # kernel, which tests/ and the decode line check against the oracle at full depth)
    parity = None
    if not args.no_parity:
        k = min(96, Tn)
        m.reset(); m.prefill(toks[:k]); lp = m.state.logits.copy()
        m.reset()
        for i in range(k):
            m.forward(int(toks[i]), i)
        ls = m.state.logits.copy()
        rel = float(np.abs(lp - ls).max() / max(float(np.abs(ls).max()), 1e-30))
        worst = rel
        if dist:
            t = torch.tensor([rel], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); worst = float(t[0].item())
        parity = {"prefix_tokens": k, "last_position_logits_maxrel_vs_token_by_token": worst, "argmax_identical": bool(int(lp.argmax()) == int(ls.argmax())),
                  "tolerance": 1e-3, "ranks": world, "ok": bool(worst < 1e-3)}
        if not parity["ok"]:
            raise SystemExit(f"prefill parity check failed: {parity}")
    for _ in range(max(args.warmup, 1)):
        m.reset(); m.prefill(toks)
    sync_all()
    sampler = ClockSampler(local_rank); sampler.start()
    ms = []
    for _ in range(args.steps):
        m.reset()
        ms.append(m.bench_prefill(toks))      # CUDA events inside the library around the prefill's kernels
    sync_all()
    clocks = sampler.stop()
    e2e = []
    for _ in range(max(1, min(args.steps, 3))):
        m.reset()
        t0 = time.perf_counter(); m.prefill(toks); e2e.append(time.perf_counter() - t0)   # tokens H2D, last logits D2H inside
    ms_total, e2e_s = float(sum(ms)), float(sum(e2e))
    if dist:
        t = torch.tensor([ms_total, e2e_s], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, e2e_s = float(t[0].item()), float(t[1].item())
    n_seq = 1 if tp > 1 else world
    ms_per_step = ms_total / args.steps
    meta = gf.meta
    kvd = meta.num_kv_heads * meta.head_dim
    layer_params = 2 * meta.embed_dim ** 2 + 2 * kvd * meta.embed_dim + 3 * meta.embed_dim * meta.interm_size
    flops = 2 * Tn * layer_params * meta.num_layers + 2 * meta.vocab_size * meta.embed_dim + meta.num_layers * 4 * meta.embed_dim * Tn * (Tn + 1) / 2
    peaks = load_peaks()
    achieved = flops / tp / (ms_per_step / 1e3) / 1e12
    out = {"metric": f"prefill tok/s ({args.dtype.upper()}, {Tn} tokens)", "value": n_seq * Tn / (ms_per_step / 1e3), "unit": "tok/s", "n_gpus": world,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if tp > 1 else "weak",
           "vs_baseline": None, "dtype": "bf16x2 split, f32 accumulate", "data": "synthetic",
           "config": {"workload": f"{tier} {T.TIERS[tier]} random-init {args.dtype.upper()}, one-pass prefill of {Tn} tokens per step", "tier": tier,
                      "dtype": args.dtype, "prefill_tokens": Tn, "parallelism": f"tp{tp}" if tp > 1 else (f"replicas x{world}" if world > 1 else "single GPU"),
                      "l2": "activations + weights exceed L2"},
           "clocks": clocks,
           "roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["tensor_sustained"], "unit": "TFLOP/s", "frac": achieved / peaks["tensor_sustained"],
                        "traffic": None, "peak_source": peaks["src"] + " bf16_tflops_sustained (kernels timed inside a long step)",
                        "useful_flops_per_step": flops, "how": "SURVEY 8d useful FLOPs (projections + causal attention + last-position LM head) / CUDA-event time"},
           "e2e": {"value": n_seq * Tn * len(e2e) / e2e_s, "unit": "tok/s", "h2d_bytes_per_step": 4 * Tn, "d2h_bytes_per_step": 4 * meta.vocab_size},
           "gpu_launches": m.launches_last_prefill * args.steps, "parity_check": parity}
    m.close()
    if rank == 0:
        print(json.dumps(out))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="decode", choices=["decode", "prefill"])
    ap.add_argument("--tier", default=None)
    ap.add_argument("--dtype", default="q4_0", choices=["q4_0", "q8_0", "f16"])
    ap.add_argument("--tokens-per-step", type=int, default=256)
    ap.add_argument("--prefill-tokens", type=int, default=2047)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--cpu-tokens", type=int, default=4)
    ap.add_argument("--no-config1", action="store_true", help="reference arm: skip BASELINE config 1 (nano Q8_0, 256 greedy tokens on the CPU port)")
    args = ap.parse_args()
    args.tier_given = args.tier is not None
    if args.tier is None:
        args.tier = "big"

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return 0
        return run_reference(args)

    from nanollama_b200 import gguf as G
    from nanollama_b200 import tiers as T
    typ = G.TYPE_IDS[args.dtype]
    metric = f"decode tok/s ({args.dtype.upper()}, bs=1)"
    config = decode_config(args, T)

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

    if args.mode == "prefill":
        rc = run_prefill(args, torch, dist, rank, world, local_rank)
        if dist:
            dist.destroy_process_group()
        return rc

    tps = args.tokens_per_step
    seq_len = min(2048, PROMPT_LEN + tps + 8)
    want_cpu = rank == 0 and world == 1 and not args.no_cpu_baseline
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
    prompt = bench_prompt(gf.meta.vocab_size)

    # ---- parity check 1: the same tier truncated to 2 layers, through the same (sharded) path, against the CPU oracle ----
    parity = {}
    if not args.no_parity:
        gf2 = T.SyntheticGGUF(args.tier, typ, seed=0, seq_len=64, layers=2)
        m2 = M.load_llama_model(gf2, device=local_rank, tp_rank=rank if tp > 1 else 0, tp_size=tp)
        worst2, ok2 = 0.0, True
        o2 = None
        if rank == 0:
            from oracle import oracle as O
            o2 = O.OracleModel(gf2)
        for pos, t in enumerate(prompt[:4]):
            m2.forward(int(t), pos)
            if o2 is not None:
                exp = o2.forward(int(t), pos)
                worst2 = max(worst2, float(np.abs(m2.state.logits - exp).max() / np.abs(exp).max()))
        got2 = m2.generate_greedy(prompt[:8], 16)
        if o2 is not None:
            exp2, mg = o2.generate_greedy(prompt[:8], 16)
            bad = [i for i in range(16) if got2[i] != exp2[i]]
            ok2 = worst2 < 1e-3 and len(got2) == len(exp2) and (not bad or float(mg[bad[0]]) < 1e-4)
            o2.close()
        m2.close()
        parity["truncated_2_layers"] = {"logits_maxrel_vs_oracle": worst2, "greedy_16_identical": bool(ok2), "tolerance": 1e-3}

    t0 = time.time()
    m = M.load_llama_model(gf, device=local_rank, tp_rank=rank if tp > 1 else 0, tp_size=tp)
    load_s = time.time() - t0
    decode_path = m.decode_path

    def sync_all():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- parity check 2: the full-depth greedy stream is the same on every rank ----
    if not args.no_parity:
        stream = m.generate_greedy(prompt, 32)
        same = True
        if dist:
            t = torch.tensor(stream.astype(np.int64), device="cuda")
            ref = t.clone()
            dist.broadcast(ref, 0)
            flag = torch.tensor([1 if torch.equal(t, ref) else 0], device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            same = bool(flag.item() == 1)
        parity["full_depth_stream_32"] = {"tokens_crc": int(np.bitwise_xor.reduce(stream.astype(np.int64) * (1 + np.arange(stream.size)))),
                                          "identical_on_all_ranks": same, "ranks": world}

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
    e2e_steps = max(1, min(args.steps, 3))
    e2e_s = sum(e2e_step() for _ in range(e2e_steps))
    if dist:
        t = torch.tensor([e2e_s], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = n_seq * tps * e2e_steps / e2e_s

    # ---- roofline of the dominant kernel (decode_tiled_kernel: one launch per token, > 99 % of the step, profiles/*launches*;
    #      models the tiled path does not take fall back to the gemv_stream_kernel chain and report that family instead) ----
    # achieved = algorithmic bytes of one decode step (weights + one embedding row + norm weights + KV read/write at
    # the mid position of the step) / CUDA-event time of that step, i.e. every launch gap, the attention and the argmax kernels are
    # charged to the GEMV too -- a lower bound of the kernel's own bandwidth.
    peaks = load_peaks()
    peak = peaks["hbm"]
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
        del raw
    except Exception as e:  # never let the extra measurement kill the bench line
        alone = {"error": str(e)}
    traffic, traffic_src = None, None
    try:   # DRAM bytes per launch of the dominant kernel from the COMMITTED ncu capture (not a measurement of this run)
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        ent = tr.get(decode_path, {}).get(f"{args.tier}/{args.dtype}")
        if ent and tp == 1:
            traffic = ent.get("dram_bytes_per_launch")
            traffic_src = f"committed ncu --set full capture {ent.get('capture')} @ git {ent.get('git')} (dram__bytes_read.sum + dram__bytes_write.sum of one launch = one token)"
    except Exception:
        traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": peaks["src"] + " hbm_gbs", "kernel": decode_path,
                "bytes_per_step": int(bytes_tok), "gemv_share_of_bytes": weight_only / bytes_tok, "kernel_alone": alone,
                "launch_us": 1000.0 * ms_per_step / tps,
                "how": "one launch of the persistent kernel = one decoded token: algorithmic bytes per token (SURVEY 8d, mid position of the "
                       "step) / CUDA-event time per token over the timed region (embedding, argmax and memset nodes, <1 % of the step, are "
                       "charged to it); kernel_alone = the same kernel as a one-phase LM-head GEMV timed back to back on L2-cold replicas"}

    config_out = dict(config)
    out = {"metric": metric, "value": value, "unit": "tok/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if tp > 1 else "weak", "vs_baseline": None, "dtype": "f32",
           "data": "synthetic", "config": config_out,
           "parallelism": (f"tp{tp} (column-split q/k/v/gate/up, row-split o/down, vocab-split LM head, in-kernel NVLink exchange x{2 * gf.meta.num_layers}/token)"
                           if tp > 1 else f"replicas x{world}" if world > 1 else "single GPU"),
           "load_s": round(load_s, 1),
           "clocks": clocks, "roofline": roofline,
           "e2e": {"value": e2e_value, "unit": "tok/s", "h2d_bytes_per_step": 8 * tps, "d2h_bytes_per_step": 4 * gf.meta.vocab_size * tps},
           "gpu_launches": m.launches_per_token * tps * args.steps, "wall_s": round(wall, 3)}

    # ---- CPU baseline leg (rank 0, N = 1 only): the port on the full-depth model, bounded sample; its logits double as the
    #      full-depth parity check of the model this line timed ----
    if want_cpu:
        eng = CpuEngine(gf)
        n = args.cpu_tokens
        toks = sample_tokens(gf.meta.vocab_size, n)
        secs, worst = [], 0.0
        m.reset()
        for k in range(3):
            s, lgs = eng.step(toks, PROMPT_LEN + k * n)
            secs.append(s)
            if not args.no_parity:
                for i, t in enumerate(toks):
                    m.forward(int(t), PROMPT_LEN + k * n + i)
                    worst = max(worst, float(np.abs(m.state.logits - lgs[i]).max() / np.abs(lgs[i]).max()))
        eng.close()
        v = n * len(secs) / float(sum(secs))
        out["cpu_baseline"] = {"value": v, "unit": "tok/s", "cores": eng.cores, "kind": "port",
                               "sample": f"full-depth {args.tier} ({gf.meta.num_layers} layers): 3 steps of {n} real Forward calls at positions >= {PROMPT_LEN} "
                                         f"({sum(secs):.1f} s of CPU work); nothing extrapolated", "step_tok_s_min_max": [n / max(secs), n / min(secs)]}
        if not args.no_parity:
            parity["full_depth_logits"] = {"positions": 3 * n, "logits_maxrel_vs_oracle": worst, "tolerance": 1e-3}
    m.close()
    if not args.no_parity:
        ok = parity.get("truncated_2_layers", {}).get("greedy_16_identical", True) and parity.get("full_depth_stream_32", {}).get("identical_on_all_ranks", True) \
            and parity.get("full_depth_logits", {}).get("logits_maxrel_vs_oracle", 0.0) < 1e-3
        parity["ok"] = bool(ok)
        out["parity_check"] = parity
    if rank == 0:
        print(json.dumps(out))
    if dist:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
