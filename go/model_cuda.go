// model_cuda.go — cgo binding that swaps the nanollama Go engine's CPU forward pass for libnanollama_cuda.so.
//
// Drop this file into the reference's go/ directory (package main) together with include/nanollama_cuda.h and the built
// libnanollama_cuda.so, delete Forward/Reset/matmulDispatch from model.go (or build with `-tags cuda` and guard them),
// and everything above the model keeps working unchanged: Engine.Generate / GenerateQuiet (main.go:152,233), the samplers
// (they read and mutate State.Logits, main.go:174-187), the REPL and serve.go.
//
// It is written against the C ABI only; it was NOT compiled in the build container (no Go toolchain there — see
// INTEGRATION.md).  tests/ drive the identical call sequence through the ctypes twin nanollama_b200/capi.py.
//
//go:build cuda

package main

/*
#cgo CFLAGS: -I${SRCDIR}/../include
#cgo LDFLAGS: -L${SRCDIR}/../nanollama_b200 -lnanollama_cuda -Wl,-rpath,${SRCDIR}/../nanollama_b200
#include <stdlib.h>
#include "nanollama_cuda.h"
*/
import "C"

import (
	"fmt"
	"runtime"
	"unsafe"
)

// cudaBackend hangs off LlamaModel (add `cuda *cudaBackend` to the struct in model.go:19-24).
type cudaBackend struct {
	h *C.nl_model
}

// nlCall runs one ABI call and, on failure, fetches its message.  nl_last_error() is thread-local on the C side and a
// goroutine may migrate between two cgo calls, so the call and the fetch run pinned to one OS thread.
func nlCall(what string, f func() C.int) error {
	runtime.LockOSThread()
	defer runtime.UnlockOSThread()
	rc := f()
	if rc == C.NL_OK {
		return nil
	}
	return fmt.Errorf("%s: libnanollama_cuda error %d: %s", what, int(rc), C.GoString(C.nl_last_error()))
}

// uploadTensor hands one GGUF tensor (raw bytes exactly as GetTensor returns them, gguf.go:561-574) to the device.
// The H2D copy has completed when the call returns, so no Go pointer is retained by C (cgo pointer rule).
func (b *cudaBackend) uploadTensor(g *GGUFFile, slot C.int, layer int, name string, optional bool) (bool, error) {
	data, info, err := g.GetTensor(name)
	if err != nil {
		if optional {
			return false, nil
		}
		return false, err
	}
	cols := int64(info.Dims[0])
	rows := int64(1)
	for d := uint32(1); d < info.NDims; d++ {
		rows *= int64(info.Dims[d])
	}
	if len(data) == 0 {
		return false, fmt.Errorf("%s: empty tensor", name)
	}
	return true, nlCall(name, func() C.int {
		return C.nl_upload_tensor(b.h, slot, C.int(layer), C.uint32_t(info.Type), C.int64_t(rows), C.int64_t(cols),
			unsafe.Pointer(&data[0]), C.size_t(len(data)))
	})
}

// LoadLlamaModelCUDA replaces LoadLlamaModel (model.go:121-174): same config defaulting, same tensor names and
// tied-embedding fallback (model.go:195-203), same wrapped errors; weights live in HBM afterwards and g.TensorData
// may be dropped.
func LoadLlamaModelCUDA(g *GGUFFile, device int) (*LlamaModel, error) {
	m := g.Meta
	cfg := C.nl_config{
		n_layers: C.int32_t(m.NumLayers), embed_dim: C.int32_t(m.EmbedDim), n_heads: C.int32_t(m.NumHeads),
		n_kv_heads: C.int32_t(m.NumKVHeads), head_dim: C.int32_t(m.HeadDim), vocab_size: C.int32_t(m.VocabSize),
		seq_len: C.int32_t(m.SeqLen), interm_size: C.int32_t(m.IntermSize),
		rms_norm_eps: C.float(m.RMSNormEps), rope_theta: C.float(m.RopeTheta),
		device: C.int32_t(device), tp_rank: 0, tp_size: 1, max_batch: 1,
	}
	if m.QKNorm {
		cfg.qk_norm = 1
	}
	if m.RopeConjugate {
		cfg.rope_conjugate = 1
	}
	b := &cudaBackend{}
	if err := nlCall("create", func() C.int { return C.nl_create(&cfg, &b.h) }); err != nil {
		return nil, err
	}
	fail := func(err error) (*LlamaModel, error) {
		C.nl_destroy(b.h)
		return nil, fmt.Errorf("load weights: %w", err)
	}
	if _, err := b.uploadTensor(g, C.NL_TOK_EMBD, -1, "token_embd.weight", false); err != nil {
		return fail(err)
	}
	if _, err := b.uploadTensor(g, C.NL_OUTPUT_NORM, -1, "output_norm.weight", false); err != nil {
		return fail(err)
	}
	if ok, err := b.uploadTensor(g, C.NL_OUTPUT, -1, "output.weight", true); err != nil {
		return fail(err)
	} else if !ok {
		fmt.Printf("[model] output.weight not found, using tied embeddings\n")
	}
	type ent struct {
		slot C.int
		name string
		opt  bool
	}
	per := []ent{
		{C.NL_ATTN_NORM, "attn_norm.weight", false}, {C.NL_FFN_NORM, "ffn_norm.weight", false},
		{C.NL_WQ, "attn_q.weight", false}, {C.NL_WK, "attn_k.weight", false}, {C.NL_WV, "attn_v.weight", false},
		{C.NL_WO, "attn_output.weight", false}, {C.NL_WGATE, "ffn_gate.weight", false}, {C.NL_WUP, "ffn_up.weight", false},
		{C.NL_WDOWN, "ffn_down.weight", false},
		{C.NL_BQ, "attn_q.bias", true}, {C.NL_BK, "attn_k.bias", true}, {C.NL_BV, "attn_v.bias", true}, {C.NL_BO, "attn_output.bias", true},
	}
	for i := 0; i < m.NumLayers; i++ {
		for _, e := range per {
			if _, err := b.uploadTensor(g, e.slot, i, fmt.Sprintf("blk.%d.%s", i, e.name), e.opt); err != nil {
				return fail(fmt.Errorf("layer %d %s: %w", i, e.name, err))
			}
		}
	}
	if err := nlCall("finalize", func() C.int { return C.nl_finalize(b.h) }); err != nil {
		C.nl_destroy(b.h)
		return nil, err
	}
	var out C.nl_config
	C.nl_get_config(b.h, &out)
	model := &LlamaModel{
		Config: LlamaConfig{
			NumLayers: int(out.n_layers), EmbedDim: int(out.embed_dim), NumHeads: int(out.n_heads), NumKVHeads: int(out.n_kv_heads),
			HeadDim: int(out.head_dim), VocabSize: int(out.vocab_size), SeqLen: int(out.seq_len), IntermSize: int(out.interm_size),
			RMSNormEps: float32(out.rms_norm_eps), RopeTheta: float32(out.rope_theta), QKNorm: out.qk_norm != 0, RopeConjugate: out.rope_conjugate != 0,
		},
		// Only the buffer the callers touch stays on the host: the samplers read and mutate it (main.go:174-187).
		State: LlamaState{Logits: make([]float32, int(out.vocab_size))},
		cuda:  b,
	}
	fmt.Printf("[model] loaded: %d layers, %d dim, %d heads, %d kv_heads, %d vocab (CUDA backend, device %d)\n",
		model.Config.NumLayers, model.Config.EmbedDim, model.Config.NumHeads, model.Config.NumKVHeads, model.Config.VocabSize, device)
	return model, nil
}

// Forward keeps the reference signature (model.go:490): no return value, result in State.Logits.
// The reference panics on an out-of-range slice index; so does this.
func (m *LlamaModel) Forward(token int, pos int) {
	err := nlCall("Forward", func() C.int {
		return C.nl_forward(m.cuda.h, C.int32_t(token), C.int32_t(pos), (*C.float)(unsafe.Pointer(&m.State.Logits[0])))
	})
	if err != nil {
		panic(err)
	}
}

// SetGammaCUDA pushes a loaded personality (gamma.go:22-35) to the device, where the embedding kernel adds row
// IndexMap[token] to the embedding exactly like ApplyToEmbedding (gamma.go:272-290, model.go:502-505); nil removes it.
// Call it wherever the reference assigns model.Gamma (main.go:79): `model.Gamma = gamma; model.SetGammaCUDA(gamma)`.
func (m *LlamaModel) SetGammaCUDA(g *GammaEssence) error {
	if g == nil || g.NumTokens == 0 {
		return nlCall("set_gamma", func() C.int { return C.nl_set_gamma(m.cuda.h, nil, 0, nil) })
	}
	if g.EmbedDim != m.Config.EmbedDim {
		return fmt.Errorf("gamma embed_dim %d != model dim %d", g.EmbedDim, m.Config.EmbedDim)
	}
	n := g.NumTokens * g.EmbedDim
	rows := make([]C.float, n)
	for i := 0; i < n; i++ {
		if g.IsF16 {
			rows[i] = C.float(half2float(g.ValuesF16[i]))
		} else {
			rows[i] = C.float(g.Values[i])
		}
	}
	t2r := make([]C.int32_t, m.Config.VocabSize)
	for i := range t2r {
		t2r[i] = -1
	}
	for tok, row := range g.IndexMap {
		if int(tok) >= 0 && int(tok) < len(t2r) {
			t2r[tok] = C.int32_t(row)
		}
	}
	return nlCall("set_gamma", func() C.int { return C.nl_set_gamma(m.cuda.h, &rows[0], C.int32_t(g.NumTokens), &t2r[0]) })
}

// Reset keeps model.go:623-631: zero both KV caches, Pos = 0.
func (m *LlamaModel) Reset() {
	if err := nlCall("Reset", func() C.int { return C.nl_reset(m.cuda.h) }); err != nil {
		panic(err)
	}
	m.State.Pos = 0
}

// GenerateGreedyCUDA is the optional fast path for `--temp 0 --rep-penalty 1.0`: the whole loop of Engine.Generate
// (main.go:152-230) stays on the device (argmax with the first-maximum tie rule of main.go:400-408).
func (m *LlamaModel) GenerateGreedyCUDA(prompt []int, maxTokens, eosID int) ([]int, error) {
	if len(prompt) == 0 {
		return nil, fmt.Errorf("generate: empty prompt")
	}
	if maxTokens <= 0 {
		return []int{}, nil
	}
	p := make([]C.int32_t, len(prompt))
	for i, t := range prompt {
		p[i] = C.int32_t(t)
	}
	out := make([]C.int32_t, maxTokens)
	var n C.int32_t
	err := nlCall("generate", func() C.int {
		return C.nl_generate_greedy(m.cuda.h, &p[0], C.int32_t(len(p)), C.int32_t(maxTokens), C.int32_t(eosID), &out[0], &n)
	})
	if err != nil {
		return nil, err
	}
	res := make([]int, int(n))
	for i := range res {
		res[i] = int(out[i])
	}
	return res, nil
}

// ForwardDeviceCUDA is Forward with the logits left on the device (State.Logits is not refreshed): the companion of SampleCUDA.
func (m *LlamaModel) ForwardDeviceCUDA(token, pos int) {
	if err := nlCall("Forward", func() C.int { return C.nl_forward(m.cuda.h, C.int32_t(token), C.int32_t(pos), nil) }); err != nil {
		panic(err)
	}
}

// SampleCUDA is one sampling step of Engine.Generate (main.go:177-197: repetition penalty over recentTokens, then
// sampleTopP when topP < 1 else sampleTopK, argmax when temp <= 0) on the device-resident logits of the last forward.
// u must be the e.rng.Float32() the host samplers would draw at this step (drawn only when temp > 0), so that a seeded
// run produces the reference's token stream.
func (m *LlamaModel) SampleCUDA(temp float32, topK int, topP, repPenalty float32, recentTokens []int, u float32) (int, error) {
	var rp *C.int32_t
	rec := make([]C.int32_t, len(recentTokens))
	for i, t := range recentTokens {
		rec[i] = C.int32_t(t)
	}
	if len(rec) > 0 {
		rp = &rec[0]
	}
	var tok C.int32_t
	err := nlCall("sample", func() C.int {
		return C.nl_sample(m.cuda.h, C.float(temp), C.int32_t(topK), C.float(topP), C.float(repPenalty), rp, C.int32_t(len(rec)), C.float(u), &tok)
	})
	if err != nil {
		return 0, err
	}
	return int(tok), nil
}

// Close releases the device memory (there is no counterpart in the reference: its weights are Go slices).
// DecodePath names the kernel family that runs a batch-1 Forward of this model ("decode_tiled_kernel", ...): for the banner / logs.
func (m *LlamaModel) DecodePath() string { return C.GoString(C.nl_decode_path(m.cuda.h)) }

func (m *LlamaModel) Close() {
	if m.cuda != nil && m.cuda.h != nil {
		C.nl_destroy(m.cuda.h)
		m.cuda.h = nil
	}
}
