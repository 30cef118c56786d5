/*
 * nanollama_cuda.h — C ABI of libnanollama_cuda.so, the B200 (sm_100a) backend for the nanollama Go engine's
 * quantized forward path.  Plain C types only; every entry point returns 0 (NL_OK) or a negative nl_status and
 * leaves a message for nl_last_error() (thread-local).  One in-flight call per nl_model (the reference engine is
 * single-threaded by contract: go/serve.go:56,106-108 holds a mutex around generation).
 *
 * Each entry point names the reference interface it replaces (paths relative to the reference repo).
 * The cgo binding a maintainer adds on the Go side is shown in INTEGRATION.md and go/model_cuda.go.
 *
 * There is no CPU fallback: every compute entry point fails with NL_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef NANOLLAMA_CUDA_H
#define NANOLLAMA_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NL_ABI_VERSION 1

typedef enum nl_status {
    NL_OK = 0,
    NL_ERR_INVALID = -1,      /* bad argument: token/pos out of range, shape mismatch, cols % 32 != 0 ...            */
    NL_ERR_CUDA = -2,         /* CUDA runtime / driver error, or no usable device                                     */
    NL_ERR_UNSUPPORTED = -3,  /* GGML tensor type the engine cannot decode (go/model.go:383-385 only prints WARNING)  */
    NL_ERR_STATE = -4,        /* call order: forward before finalize, upload after finalize, missing tensor ...       */
    NL_ERR_OOM = -5
} nl_status;

/* GGML tensor type ids, go/gguf.go:44-56 */
enum { NL_F32 = 0, NL_F16 = 1, NL_Q4_0 = 2, NL_Q5_0 = 6, NL_Q8_0 = 8, NL_Q4_K = 12, NL_Q6_K = 14 };

/* Mirrors LlamaConfig, go/model.go:27-42 (same meaning, same defaults applied by the caller: HeadDim =
 * EmbedDim/NumHeads when 0 (model.go:140), SeqLen capped to 2048 (model.go:145-148) — nl_create applies both). */
typedef struct nl_config {
    int32_t n_layers;        /* NumLayers   */
    int32_t embed_dim;       /* EmbedDim    */
    int32_t n_heads;         /* NumHeads    */
    int32_t n_kv_heads;      /* NumKVHeads  */
    int32_t head_dim;        /* HeadDim     */
    int32_t vocab_size;      /* VocabSize   */
    int32_t seq_len;         /* SeqLen      */
    int32_t interm_size;     /* IntermSize  */
    float rms_norm_eps;      /* RMSNormEps  */
    float rope_theta;        /* RopeTheta   */
    int32_t qk_norm;         /* QKNorm        (bool) */
    int32_t rope_conjugate;  /* RopeConjugate (bool) */
    /* --- backend placement (no counterpart in the reference, which is single-process CPU) --- */
    int32_t device;          /* CUDA device ordinal for this process                                      */
    int32_t tp_rank;         /* tensor-parallel rank of this process (0 when tp_size == 1)                */
    int32_t tp_size;         /* 1, 2, 4 or 8: column-split q/k/v/gate/up, row-split o/down, vocab-split LM head */
    int32_t max_batch;       /* independent sequences decoded per step by nl_forward_batch (>=1)          */
} nl_config;

/* Tensor slots = the fields of LlamaWeights / LlamaLayerWeights, go/model.go:45-90, filled by loadWeights
 * (go/model.go:177-265) from GGUF names token_embd / output_norm / output / blk.N.{attn_norm,ffn_norm,attn_q,
 * attn_k,attn_v,attn_output,ffn_gate,ffn_up,ffn_down}.weight and the optional blk.N.attn_{q,k,v,output}.bias */
typedef enum nl_slot {
    NL_TOK_EMBD = 0, NL_OUTPUT_NORM = 1, NL_OUTPUT = 2,
    NL_ATTN_NORM = 3, NL_FFN_NORM = 4, NL_WQ = 5, NL_WK = 6, NL_WV = 7, NL_WO = 8, NL_WGATE = 9, NL_WUP = 10, NL_WDOWN = 11,
    NL_BQ = 12, NL_BK = 13, NL_BV = 14, NL_BO = 15,
    NL_SLOT_COUNT = 16
} nl_slot;

typedef struct nl_model nl_model;

/* message of the last failing call on this thread ("" if none) */
const char *nl_last_error(void);
int nl_abi_version(void);
/* number of visible CUDA devices with compute capability 10.x (0 => every compute call will fail) */
int nl_device_count(void);

/* ---- model lifetime: replaces LoadLlamaModel / loadWeights / allocState / precomputeRoPE, go/model.go:121-358 ---- */
int nl_create(const nl_config *cfg, nl_model **out);
/* Copies the FULL (unsharded) tensor exactly as it sits in the GGUF data blob (GGUFFile.GetTensor, go/gguf.go:561-574).
 * The H2D copy has completed when the call returns (cgo may not retain Go pointers); with tp_size > 1 the library keeps
 * only this rank's shard.  rows x cols are the engine's [out_features, in_features]; vectors use rows = 1.
 * layer is ignored for the three global slots.  Not uploading NL_OUTPUT selects the tied-embedding fallback
 * (go/model.go:195-203). */
int nl_upload_tensor(nl_model *m, int slot, int layer, uint32_t ggml_type, int64_t rows, int64_t cols,
                     const void *host_data, size_t nbytes);
/* Optional gamma essence (go/model.go:503-505, go/gamma.go:272-290): dense fp32 rows [n_rows, embed_dim] and a
 * token -> row map (-1 = untouched) of length vocab_size.  Pass rows = NULL to clear. */
int nl_set_gamma(nl_model *m, const float *rows, int32_t n_rows, const int32_t *token_to_row);
/* Validates that every required tensor is present, builds RoPE tables (float64 pow/cos/sin -> fp32, go/model.go:346-358),
 * allocates the fp32 KV cache [layer][seq][kv_dim] (go/model.go:337-338) and captures the per-token CUDA graph. */
int nl_finalize(nl_model *m);
void nl_destroy(nl_model *m);
/* the config after nl_create's defaulting (HeadDim, SeqLen cap) */
int nl_get_config(const nl_model *m, nl_config *out);

/* ---- (*LlamaModel).Forward(token, pos), go/model.go:490-620.  logits_out: caller-owned host buffer of vocab_size floats
 * (= State.Logits), or NULL to leave them on the device. ---- */
int nl_forward(nl_model *m, int32_t token, int32_t pos, float *logits_out);
/* ---- (*LlamaModel).Reset(), go/model.go:623-631: zero both KV caches ---- */
int nl_reset(nl_model *m);
/* copy the device-resident logits of the last forward to the host (State.Logits read, go/main.go:174) */
int nl_get_logits(nl_model *m, float *logits_out);

/* ---- Engine.Generate's loop with temp <= 0 and rep-penalty 1.0 (go/main.go:152-230, argmax :400-408), kept on the
 * device: Reset, token-by-token prefill (stops at seq_len-1), then argmax -> Forward until n_new tokens, EOS
 * (eos_id < 0 disables) or pos reaches seq_len.  out_tokens: n_new int32; *n_out = tokens produced (a terminal EOS
 * is included, like the reference's token count). ---- */
int nl_generate_greedy(nl_model *m, const int32_t *prompt, int32_t n_prompt, int32_t n_new, int32_t eos_id,
                       int32_t *out_tokens, int32_t *n_out);

/* ---- one sampling step of Engine.Generate (go/main.go:177-197, :294-398) on the logits the last nl_forward left on the device,
 * so that they never cross PCIe and the host does not sort the vocabulary: repetition penalty over `recent` (n_recent tokens, once
 * per occurrence, applied in place like the reference does to State.Logits), then sampleTopP when top_p < 1, else sampleTopK;
 * temperature <= 0: argmax.  u = the uniform number rng.Float32() the reference would draw at this step (drawn by the host, so the
 * random stream stays the host's; ignored when temperature <= 0).  *token_out = the sampled token id. ---- */
int nl_sample(nl_model *m, float temperature, int32_t top_k, float top_p, float rep_penalty, const int32_t *recent, int32_t n_recent,
              float u, int32_t *token_out);

/* ---- B independent sequences, one token each (small-batch decode; same weights, per-sequence KV cache).
 * tokens/pos: B int32 each; logits_out: [B, vocab_size] host floats or NULL.  B <= max_batch. ---- */
int nl_forward_batch(nl_model *m, int32_t B, const int32_t *tokens, const int32_t *pos, float *logits_out);

/* ---- prefill of n prompt tokens at positions pos0..pos0+n-1 of sequence 0 in one pass (the reference feeds them one
 * Forward at a time, go/main.go:160-166); logits_last = logits after the final token, host buffer or NULL ---- */
int nl_prefill(nl_model *m, const int32_t *tokens, int32_t n, int32_t pos0, float *logits_last);

/* ---- operator-level hooks (parity + microbench) ---- */
/* Dequant* of go/quant.go:18-42,100-118,405-430,296-333,174-208 and half2float (go/gguf.go:634): n elements of
 * ggml_type at host_src -> fp32 at host_dst, computed on the device from the library's resident weight layout. */
int nl_dequant(uint32_t ggml_type, const void *host_src, int64_t n, float *host_dst);
/* matmulDispatch(out, w, wtype, x, rows, cols), go/model.go:361-386, host buffers in and out (x: [batch, cols],
 * out: [batch, rows]; batch = 1 is the reference's case). */
int nl_matmul(uint32_t ggml_type, const void *host_w, int64_t rows, int64_t cols, const float *host_x, int32_t batch,
              float *host_out);

/* Device-resident matrix for repeated matmuls (weights uploaded once, like model weights). */
typedef struct nl_matrix nl_matrix;
int nl_matrix_create(uint32_t ggml_type, const void *host_w, int64_t rows, int64_t cols, int32_t device, nl_matrix **out);
int nl_matrix_matmul(nl_matrix *w, const float *host_x, int32_t batch, float *host_out);
/* times `iters` back-to-back device-resident GEMVs with CUDA events (after `warmup`), x already in HBM; rotates over
 * `n_copies` replicas of the matrix so successive launches do not hit L2 (n_copies*bytes > L2). ms_out = mean ms/launch */
int nl_matrix_bench(nl_matrix *w, int32_t batch, int32_t n_copies, int32_t warmup, int32_t iters, float *ms_out);
void nl_matrix_destroy(nl_matrix *w);

/* ---- measurement hooks ---- */
/* Runs n_steps decode steps (device-side greedy feedback, starting from `token` at `pos0`) and returns the CUDA-event
 * time of the whole span in ms.  No host<->device traffic inside the timed region. */
int nl_bench_decode(nl_model *m, int32_t token, int32_t pos0, int32_t n_steps, float *ms_out);
/* nl_prefill of device-resident work only: the n tokens are copied to the device first, then the prefill's kernels are timed with
 * CUDA events on the library's stream (ms_out); no logits leave the device.  *launches_out (optional) = kernels launched. */
int nl_bench_prefill(nl_model *m, const int32_t *tokens, int32_t n, int32_t pos0, float *ms_out, int32_t *launches_out);
/* number of kernel launches one decode step issues (graph nodes) */
int nl_launches_per_token(const nl_model *m);
/* bytes of weights resident on this device (after sharding) */
int64_t nl_weight_bytes(const nl_model *m);
/* name of the kernel family that executes a batch-1 Forward of this model ("decode_tiled_kernel",
 * "gemv_stream_kernel chain"); static string */
const char *nl_decode_path(const nl_model *m);

/* ---- tensor-parallel plumbing (tp_size > 1): peers exchange 64-byte CUDA IPC handles of their all-reduce windows
 * through the host's process group (torch.distributed in the Python host), then hand them back here. ---- */
int nl_tp_export_handle(nl_model *m, void *handle64);
int nl_tp_import_handles(nl_model *m, const void *handles64_by_rank, int32_t n_ranks);

#ifdef __cplusplus
}
#endif
#endif /* NANOLLAMA_CUDA_H */
