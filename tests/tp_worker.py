"""torchrun worker for the tensor-parallel parity test: every rank loads the same GGUF with tp_size = WORLD_SIZE, rank 0 also
runs the CPU oracle; logits (tolerance 1e-3 max-rel, BASELINE.json) and the greedy stream must match on every rank."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from nanollama_b200 import gguf as G
    from nanollama_b200 import model as M
    from nanollama_b200 import tiers as T
    from oracle import oracle as O

    from tests.helpers import WithBias
    cases = [("tiny_mha_qknorm_q8_0", G.load_gguf(os.path.join(ROOT, "tests", "golden", "tiny_mha_qknorm_q8_0.gguf")), 40),
             ("goldie-4L-q4_0", T.SyntheticGGUF("goldie", G.GGML_Q4_0, seed=2, seq_len=128, layers=4, vocab=4096), 24)]
    cases.append(("big-2L-q4_0", T.SyntheticGGUF("big", G.GGML_Q4_0, seed=4, seq_len=64, layers=2, vocab=4096), 8))
    # all four optional bias tensors under tensor parallelism (q/k/v biases are sharded with their rows, the o bias is added by rank 0)
    cases.append(("goldie-2L-q4_0+bias", WithBias(T.SyntheticGGUF("goldie", G.GGML_Q4_0, seed=6, seq_len=96, layers=2, vocab=4096), seed=2), 16))
    if world <= 4:
        cases.append(("large-2L-q8_0", T.SyntheticGGUF("large", G.GGML_Q8_0, seed=3, seq_len=64, layers=2, vocab=4096), 8))
    ok = True
    for name, gf, n_new in cases:
        try:
            from nanollama_b200.tp import shard_plan
            shard_plan(gf.meta, world)
        except ValueError as e:
            if rank == 0:
                print(f"[tp] {name}: skipped ({e})")
            continue
        m = M.load_llama_model(gf, device=local, tp_rank=rank, tp_size=world)
        rng = np.random.default_rng(5)
        prompt = np.concatenate([[1], rng.integers(3, gf.meta.vocab_size, size=7)]).astype(np.int32)
        worst = 0.0
        o = O.OracleModel(gf) if rank == 0 else None
        m.reset()
        for pos, t in enumerate(prompt):
            m.forward(int(t), pos)
            if o is not None:
                exp = o.forward(int(t), pos)
                worst = max(worst, float(np.abs(m.state.logits - exp).max() / np.abs(exp).max()))
        # one-pass prefill on the tensor cores under tensor parallelism (row-split GEMMs + reduce-scatter / all-gather of the rows):
        # last-position logits of a 40-token prompt against the oracle, then decode continues from the prefilled cache
        if gf.meta.seq_len >= 48:
            long_prompt = np.concatenate([[1], rng.integers(3, gf.meta.vocab_size, size=39)]).astype(np.int32)
            m.reset()
            m.prefill(long_prompt)
            pf = m.state.logits.copy()
            nxt = int(long_prompt[3])
            m.forward(nxt, len(long_prompt))
            if o is not None:
                o.reset()
                for pos, t in enumerate(long_prompt):
                    exp = o.forward(int(t), pos)
                w1 = float(np.abs(pf - exp).max() / np.abs(exp).max())
                exp = o.forward(nxt, len(long_prompt))
                w2 = float(np.abs(m.state.logits - exp).max() / np.abs(exp).max())
                print(f"[tp] {name} tp={world}: prefill(40) logits max-rel {w1:.2e}, next decode step {w2:.2e}")
                ok = ok and w1 < 1e-3 and w2 < 1e-3
        got = m.generate_greedy(prompt, n_new)
        # every rank must hold the same stream (full logits are gathered into every window)
        t = torch.tensor(got.astype(np.int64), device="cuda")
        ref = t.clone()
        dist.broadcast(ref, 0)
        same = bool(torch.equal(t, ref))
        if o is not None:
            exp, margins = o.generate_greedy(prompt, n_new)
            bad = [i for i in range(min(len(exp), len(got))) if got[i] != exp[i]]
            stream_ok = len(got) == len(exp) and (not bad or margins[bad[0]] < 1e-4)
            print(f"[tp] {name} tp={world}: logits max-rel {worst:.2e}, greedy {'identical' if not bad else 'diverges at %d (margin %.2e)' % (bad[0], margins[bad[0]])}")
            ok = ok and worst < 1e-3 and stream_ok
        ok = ok and same
        m.close()
        dist.barrier()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if rank == 0:
        print("TP_PARITY_OK" if flag.item() == 1 else "TP_PARITY_FAILED")
    return 0 if flag.item() == 1 else 1


if __name__ == "__main__":
    sys.exit(main())
