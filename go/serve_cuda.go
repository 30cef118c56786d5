// serve_cuda.go — continuous batching behind /chat for the CUDA backend (SURVEY.md 8f row 3).
//
// The reference serialises requests with one mutex around engine.GenerateQuiet (serve.go:56, :106-108) because its model holds one
// sequence.  The CUDA model holds max_batch sequences (per-sequence KV caches, nl_forward_batch), so the mutex becomes this
// scheduler: every step is ONE forward over all sequences in flight; requests join at the next step boundary and leave when they
// finish; each request still runs exactly the loop of GenerateQuiet (main.go:233-291): same prompt feeding, repetition penalty,
// samplers (sampleTopK / sampleTopP on that sequence's logits row) and stop rules.
//
// Drop-in: in runServer replace `mu.Lock(); result := engine.GenerateQuiet(prompt, params); mu.Unlock()` by
// `result := batcher.Generate(prompt, params)`.  Written against the C ABI only and NOT compiled in the build container (no Go
// toolchain, see INTEGRATION.md); nanollama_b200/batcher.py is the same scheduler, line for line, and is what tests/ exercise
// (tests/test_batcher_host.py on a fake model, tests/test_gpu_parity.py::test_continuous_batcher_... on the GPU).
//
//go:build cuda

package main

/*
#include "nanollama_cuda.h"
*/
import "C"

import (
	"math/rand"
	"sync"
	"unsafe"
)

type batchSeq struct {
	prompt    []int
	params    GenParams
	rng       *rand.Rand
	pos, fed  int
	next      int
	recent    []int
	out       []byte
	steps     int
	done      chan string
}

// Batcher owns the model: one goroutine (run) issues every forward.
type Batcher struct {
	engine  *Engine
	B       int
	slots   []*batchSeq
	mu      sync.Mutex
	waiting []*batchSeq
	wake    chan struct{}
	logits  []float32 // [B][vocab], filled by nl_forward_batch
}

func NewBatcher(e *Engine, maxBatch int) *Batcher {
	b := &Batcher{engine: e, B: maxBatch, slots: make([]*batchSeq, maxBatch), wake: make(chan struct{}, 1),
		logits: make([]float32, maxBatch*e.model.Config.VocabSize)}
	go b.run()
	return b
}

// Generate is GenerateQuiet for one request; it blocks until that request's text is complete.
func (b *Batcher) Generate(prompt string, p GenParams) string {
	tokens := b.engine.tokenizer.Encode(prompt, true)
	if len(tokens) == 0 {
		return ""
	}
	s := &batchSeq{prompt: tokens, params: p, rng: rand.New(rand.NewSource(rand.Int63())), next: tokens[0], done: make(chan string, 1)}
	b.mu.Lock()
	b.waiting = append(b.waiting, s)
	b.mu.Unlock()
	select {
	case b.wake <- struct{}{}:
	default:
	}
	return <-s.done
}

func (b *Batcher) run() {
	for {
		if !b.step() {
			<-b.wake
		}
	}
}

// step: admit waiting requests into free slots, run one batch forward over rows 0..highest busy slot (idle rows carry a dummy
// token at position 0 that nothing reads: a slot is refilled from position 0 when it is reused), advance every sequence.
func (b *Batcher) step() bool {
	b.mu.Lock()
	for i := range b.slots {
		if b.slots[i] == nil && len(b.waiting) > 0 {
			b.slots[i] = b.waiting[0]
			b.waiting = b.waiting[1:]
		}
	}
	b.mu.Unlock()
	n := 0
	for i, s := range b.slots {
		if s != nil {
			n = i + 1
		}
	}
	if n == 0 {
		return false
	}
	toks := make([]C.int32_t, n)
	pos := make([]C.int32_t, n)
	for i := 0; i < n; i++ {
		if s := b.slots[i]; s != nil {
			toks[i], pos[i] = C.int32_t(s.next), C.int32_t(s.pos)
		}
	}
	m := b.engine.model
	err := nlCall("forward_batch", func() C.int {
		return C.nl_forward_batch(m.cuda.h, C.int32_t(n), &toks[0], &pos[0], (*C.float)(unsafe.Pointer(&b.logits[0])))
	})
	if err != nil {
		panic(err)
	}
	vocab := m.Config.VocabSize
	for i := 0; i < n; i++ {
		if b.slots[i] != nil {
			b.afterForward(i, b.logits[i*vocab:(i+1)*vocab])
		}
	}
	return true
}

// sampleTopKWith / sampleTopPWith are sampleTopK / sampleTopP (main.go:294-343, :346-398) with the random source passed in instead of
// read from e.rng -- a two-line refactor of the reference functions (each request keeps its own stream), not shown here.

func (b *Batcher) finish(i int) {
	s := b.slots[i]
	b.slots[i] = nil
	s.done <- string(s.out)
}

// afterForward: what GenerateQuiet does between two Forward calls (main.go:240-288), for the sequence in slot i.
func (b *Batcher) afterForward(i int, logits []float32) {
	s, e := b.slots[i], b.engine
	cfg := e.model.Config
	s.pos++
	if s.fed < len(s.prompt) {
		s.fed++
		if s.fed < len(s.prompt) && s.pos < cfg.SeqLen-1 { // prompt feeding stops at seq_len-1 (main.go:244)
			s.next = s.prompt[s.fed]
			return
		}
		s.fed = len(s.prompt)
	} else if s.pos >= cfg.SeqLen { // main.go:286-288
		b.finish(i)
		return
	}
	if s.steps >= s.params.MaxTokens || len(s.out) >= 8192 { // main.go:252
		b.finish(i)
		return
	}
	s.steps++
	if e.repPenalty > 1.0 { // main.go:253-263, in place on this row
		for _, tok := range s.recent {
			if tok >= 0 && tok < cfg.VocabSize {
				if logits[tok] > 0 {
					logits[tok] /= e.repPenalty
				} else {
					logits[tok] *= e.repPenalty
				}
			}
		}
	}
	var next int
	if s.params.TopP < 1.0 {
		next = sampleTopPWith(s.rng, logits, cfg.VocabSize, s.params.Temperature, s.params.TopP) // sampleTopP (main.go:346-398) with the request's own rng
	} else {
		next = sampleTopKWith(s.rng, logits, cfg.VocabSize, s.params.Temperature, s.params.TopK) // sampleTopK (main.go:294-343)
	}
	s.recent = append(s.recent, next)
	if len(s.recent) > e.repWindow {
		s.recent = s.recent[1:]
	}
	if next == e.tokenizer.EosID {
		b.finish(i)
		return
	}
	s.out = append(s.out, e.tokenizer.DecodeToken(next)...)
	if s.steps >= s.params.MaxTokens { // the reference still runs Forward for the last sampled token; nothing reads its logits
		b.finish(i)
		return
	}
	s.next = next
}
