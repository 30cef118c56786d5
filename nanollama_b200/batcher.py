"""Continuous batching behind /chat: the reference serialises requests with one global mutex around
Engine.GenerateQuiet (go/serve.go:56, :106-108) because its model holds ONE sequence; this backend holds ``max_batch`` sequences
(per-sequence KV caches, nl_forward_batch), so the mutex becomes a batcher: every step is ONE forward over all sequences in flight,
requests join at the next step boundary and leave when they finish, and each request still sees exactly the loop of
GenerateQuiet (go/main.go:233-291) -- same prompt feeding, repetition penalty, samplers, stop rules, its own RNG stream.

    b = ContinuousBatcher(model)                      # model.max_batch sequences
    t = b.submit(prompt_tokens, GenParams(...), seed=7)
    b.run_until_idle()                                # or b.start() for a worker thread, as the HTTP handlers would use it
    t.result()                                        # the tokens Engine(model, seed=7).generate_tokens(prompt_tokens, params) returns

A sequence occupies one slot (= batch row = KV cache) from admission to its last token; a step runs rows 0 .. highest busy slot and
feeds idle rows in between a dummy (token 0, position 0) that nothing reads (a slot is refilled from position 0 when it is reused).
go/serve_cuda.go shows the same scheduler on the Go side of the C ABI.
"""
from __future__ import annotations

import random
import threading
from collections import deque
from typing import Callable, Deque, List, Optional, Sequence

import numpy as np

from .engine import GenParams, sample_top_k, sample_top_p


class Ticket:
    """What submit() hands back: wait() / result()."""

    def __init__(self):
        self._done = threading.Event()
        self.tokens: List[int] = []
        self.error: Optional[BaseException] = None

    def wait(self, timeout: Optional[float] = None) -> bool:
        return self._done.wait(timeout)

    def result(self, timeout: Optional[float] = None) -> List[int]:
        if not self._done.wait(timeout):
            raise TimeoutError("request still in flight")
        if self.error is not None:
            raise self.error
        return self.tokens


class _Seq:
    """One request's GenerateQuiet state (go/main.go:233-291)."""

    def __init__(self, prompt: Sequence[int], p: GenParams, seed, ticket: Ticket, on_token):
        self.prompt = [int(t) for t in prompt]
        self.p = p
        self.rng = random.Random(seed)
        self.ticket = ticket
        self.on_token = on_token
        self.pos = 0            # next position to be written
        self.fed = 0            # prompt tokens fed so far
        self.next_token = None  # token the next forward carries
        self.recent: List[int] = []
        self.out_bytes = 0
        self.steps = 0          # sampling steps taken


class ContinuousBatcher:
    def __init__(self, model, eos_id: int = 2, rep_penalty: float = 1.15, rep_window: int = 64,
                 decode_token: Optional[Callable[[int], str]] = None):
        self.model = model
        self.B = int(model.max_batch)
        self.eos_id, self.rep_penalty, self.rep_window = eos_id, float(rep_penalty), int(rep_window)
        self.decode_token = decode_token or (lambda t: "")
        self.slots: List[Optional[_Seq]] = [None] * self.B
        self.waiting: Deque[_Seq] = deque()
        self.lock = threading.Lock()
        self.wake = threading.Condition(self.lock)
        self.steps_run = 0          # batch forwards issued
        self.rows_run = 0           # useful rows over all of them (occupancy = rows_run / steps_run)
        self._thread: Optional[threading.Thread] = None
        self._stop = False

    # ---- request side (any thread) ----
    def submit(self, prompt_tokens: Sequence[int], params: GenParams, seed=None, on_token=None) -> Ticket:
        t = Ticket()
        if len(prompt_tokens) == 0:
            t.error = ValueError("empty prompt")
            t._done.set()
            return t
        with self.wake:
            self.waiting.append(_Seq(prompt_tokens, params, seed, t, on_token))
            self.wake.notify()
        return t

    # ---- scheduler side (one thread) ----
    def _admit(self):
        with self.lock:
            for i in range(self.B):
                if self.slots[i] is None and self.waiting:
                    s = self.waiting.popleft()
                    s.next_token = s.prompt[0]
                    self.slots[i] = s

    def _finish(self, i: int):
        s = self.slots[i]
        self.slots[i] = None
        s.ticket._done.set()

    def _after_forward(self, i: int, logits: np.ndarray):
        """The part of GenerateQuiet between two Forward calls, for the sequence in slot i; sets its next token or retires it."""
        s, vocab, seq_len = self.slots[i], self.model.config.vocab_size, self.model.config.seq_len
        s.pos += 1
        if s.fed < len(s.prompt):
            s.fed += 1
            if s.fed < len(s.prompt) and s.pos < seq_len - 1:      # prompt feeding stops at seq_len - 1 (main.go:240-246)
                s.next_token = s.prompt[s.fed]
                return
            s.fed = len(s.prompt)
        elif s.pos >= seq_len:                                     # Forward at the last position done: the loop breaks (main.go:286-288)
            self._finish(i)
            return
        if s.steps >= s.p.max_tokens or s.out_bytes >= 8192:       # loop bounds (main.go:252)
            self._finish(i)
            return
        s.steps += 1
        if self.rep_penalty > 1.0 and s.recent:                    # main.go:253-263 -- in place
            pen = np.float32(self.rep_penalty)
            for tok in s.recent:
                if 0 <= tok < vocab:
                    logits[tok] = np.float32(logits[tok] / pen) if logits[tok] > 0 else np.float32(logits[tok] * pen)
        nxt = sample_top_p(logits, vocab, s.p.temperature, s.p.top_p, s.rng) if s.p.top_p < 1.0 else sample_top_k(logits, vocab, s.p.temperature, s.p.top_k, s.rng)
        s.recent.append(nxt)
        if len(s.recent) > self.rep_window:
            s.recent = s.recent[1:]
        if nxt == self.eos_id:
            self._finish(i)
            return
        s.ticket.tokens.append(nxt)
        s.out_bytes += len(self.decode_token(nxt).encode("utf-8"))
        if s.on_token:
            s.on_token(nxt)
        if s.steps >= s.p.max_tokens:      # the reference still runs Forward for the last sampled token; nothing reads its logits
            self._finish(i)
            return
        s.next_token = nxt

    def step(self) -> bool:
        """Admit waiting requests, run one batch forward, advance every sequence.  False when nothing is in flight."""
        self._admit()
        busy = [i for i in range(self.B) if self.slots[i] is not None]
        if not busy:
            return False
        n = busy[-1] + 1
        toks = [self.slots[i].next_token if self.slots[i] is not None else 0 for i in range(n)]
        pos = [self.slots[i].pos if self.slots[i] is not None else 0 for i in range(n)]
        try:
            logits = self.model.forward_batch(toks, pos) if n > 1 else self._forward_one(toks[0], pos[0])
        except BaseException as e:   # a failing forward fails every request in flight, not the scheduler
            for i in busy:
                self.slots[i].ticket.error = e
                self._finish(i)
            return True
        self.steps_run += 1
        self.rows_run += len(busy)
        for i in busy:
            self._after_forward(i, logits[i])
        return True

    def _forward_one(self, token: int, pos: int) -> np.ndarray:
        self.model.forward(token, pos)
        return self.model.state.logits.reshape(1, -1)

    def run_until_idle(self):
        while self.step() or self.waiting:
            pass

    # ---- worker thread for a server ----
    def start(self):
        def loop():
            while True:
                with self.wake:
                    while not self._stop and not self.waiting and all(s is None for s in self.slots):
                        self.wake.wait()
                    if self._stop:
                        return
                self.step()
        self._thread = threading.Thread(target=loop, daemon=True)
        self._thread.start()

    def stop(self):
        with self.wake:
            self._stop = True
            self.wake.notify_all()
        if self._thread:
            self._thread.join()
