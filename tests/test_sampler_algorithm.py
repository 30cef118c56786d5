"""The algorithm of csrc/nl_sample.cu restated in numpy and checked against the host mirror of go/main.go:294-398 (engine.py): candidate
selection by the top 12 bits of an order-preserving key, stable compaction, stable radix order, whole-vocabulary fallback, and the
reference's sequential fp32 sums.  The GPU test (test_gpu_parity.py::test_device_sampling_matches_host_samplers) checks the kernel
itself; this one pins the design on the CPU, including the corner cases a random model rarely produces (ties, flat and one-hot
distributions, negative zero, candidates outgrown by the nucleus)."""
from types import SimpleNamespace

import numpy as np
import pytest

from nanollama_b200.engine import Engine

HBITS, CAND_MAX, CAND_TOPP = 12, 16384, 2048   # SP_HBITS, SP_CAND_MAX, SP_CAND_TOPP of nl_sample.cu


def desc_key(lg):
    b = lg.view(np.uint32)
    asc = np.where(b & np.uint32(0x80000000), ~b, b | np.uint32(0x80000000)).astype(np.uint32)
    return (~asc).astype(np.uint32)


def radix_order(keys, idx):
    for p in range(8):   # a pass with per-thread contiguous chunks and a [digit][thread] scan IS a stable sort by that digit
        o = np.argsort((keys >> np.uint32(4 * p)) & np.uint32(15), kind="stable")
        keys, idx = keys[o], idx[o]
    return idx


def soft(x, mx, temp):
    return np.exp(((x - mx) / np.float32(temp)).astype(np.float32).astype(np.float64)).astype(np.float32)


def kernel_algorithm(lg, temp, top_k, top_p, u, stats):
    n = lg.size
    if not temp > 0:
        return int(np.argmax(lg))
    mx = lg.max()
    topp = top_p < 1.0
    inv = np.float32(1.0)
    if topp:
        inv = np.float32(1.0) / np.cumsum(soft(lg, mx, temp), dtype=np.float32)[-1]   # sequential fp32 sum in index order
    want = min(n, CAND_TOPP) if topp else min(n, top_k)
    keys = desc_key(lg)
    for attempt in ((1,) if want > CAND_MAX else (0, 1)):
        if attempt == 0:
            bins = (keys >> np.uint32(32 - HBITS)).astype(np.int64)
            pre = np.cumsum(np.bincount(bins, minlength=1 << HBITS))
            bstar = int(np.argmax(pre >= want))
            if pre[bstar] > CAND_MAX:
                continue
            sel = np.nonzero(bins <= bstar)[0].astype(np.int32)
            order = radix_order(keys[sel], sel)
        else:
            order = radix_order(keys.copy(), np.arange(n, dtype=np.int32))
        stats[attempt] += 1
        limit = order.size if topp else min(order.size, top_k)
        w = soft(lg[order[:limit]], mx, temp)
        if topp:
            w = (w * inv).astype(np.float32)
        cum = np.cumsum(w, dtype=np.float32)
        if topp:
            hit = np.nonzero(cum >= np.float32(top_p))[0]
            if hit.size == 0:
                if attempt == 0 and order.size < n:
                    continue
                return int(order[0])
            found = int(hit[0])
        else:
            found = limit - 1
        r = np.float32(u) * cum[found]
        pick = np.nonzero(r <= cum[: found + 1])[0]
        return int(order[pick[0]] if pick.size else order[0])
    raise AssertionError("unreachable")


def logits_case(name, n, rng):
    if name == "peaked":
        lg = (rng.standard_normal(n) * 4).astype(np.float32)
    elif name == "flat":
        lg = np.full(n, 0.5, np.float32)
    elif name == "near_flat":
        lg = (rng.standard_normal(n) * 0.02).astype(np.float32)
    elif name == "one_hot":
        lg = np.full(n, -30.0, np.float32); lg[n // 3] = 12.0
    elif name == "ties_and_zeros":
        lg = rng.integers(-3, 4, size=n).astype(np.float32); lg[::7] = -0.0; lg[1::7] = 0.0
    else:
        raise ValueError(name)
    return lg


@pytest.mark.parametrize("name,n", [("peaked", 32000), ("flat", 5000), ("near_flat", 40000), ("one_hot", 4096), ("ties_and_zeros", 3000), ("peaked", 300)])
def test_kernel_algorithm_selects_the_reference_token(name, n):
    import zlib
    rng = np.random.default_rng(zlib.crc32(name.encode()) % 1000 + n)
    lg = logits_case(name, n, rng)
    model = SimpleNamespace(state=SimpleNamespace(logits=lg.copy()), config=SimpleNamespace(vocab_size=n))
    host = Engine(model, seed=0)
    stats = {0: 0, 1: 0}
    for temp, top_k, top_p in [(0.8, 50, 0.9), (1.0, 40, 1.0), (0.7, 1, 1.0), (1.3, 20000, 1.0), (0.9, 50, 0.5), (3.0, 50, 0.999), (0.0, 50, 0.9)]:
        for u in (0.0, 0.2, 0.61, 0.97, float(np.nextafter(np.float32(1), np.float32(0)))):
            host.rng = SimpleNamespace(random=lambda: u)
            exp = host.sample_top_p(temp, top_p) if top_p < 1.0 else host.sample_top_k(temp, top_k)
            got = kernel_algorithm(lg, temp, top_k, top_p, u, stats)
            if got != exp:
                # Only the order among EQUAL entries may differ: -0.0 sorts below +0.0 by key but compares equal in the reference, and
                # sampleTopP orders by NORMALISED probability, where distinct logits can round to the same fp32 value (the kernel orders
                # by logit; the reference's sort.Slice is not stable, so the order among equal probabilities is arbitrary there too)
                assert temp > 0 and name in ("ties_and_zeros", "near_flat"), (name, temp, top_k, top_p, u)
                w = soft(lg, lg.max(), temp)
                if top_p < 1.0:
                    w = (w * (np.float32(1.0) / np.cumsum(w, dtype=np.float32)[-1])).astype(np.float32)
                assert w[got] == w[exp], (name, temp, top_k, top_p, u)
    if name == "near_flat":
        assert stats[1] > 0   # more tokens share the boundary bin than SP_CAND_MAX: the whole-vocabulary fallback was exercised
    if name == "peaked":
        assert stats[0] > 0   # the candidate path served the ordinary cases (top_k = 20000 and top_p = 0.999 at temp 3 fall back)
