// nl_api.cu — model object, per-token launch sequence (CUDA graph) and the C ABI of libnanollama_cuda.so.
// Replaces LoadLlamaModel / Forward / Reset (go/model.go:121-631) and the greedy loop of Engine.Generate (go/main.go:152-230).
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <vector>

#include <stdlib.h>

#include "nl_gemv.cuh"
#include "nl_kernels.cuh"
#include "nl_stream.cuh"
#include "nl_tile.cuh"
#include "nl_sample.cuh"
#include <string>
#include "nl_tp.cuh"
#include "nl_gemm.cuh"
#include "nl_gemm2.cuh"
#include "nl_prefill.cuh"

namespace nl {

static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap);
}
int fail(int code, const char *fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap);
    return code;
}

static bool type_planar(int t) { return t == NL_Q4_0 || t == NL_Q8_0; }
static bool type_supported(int t) { return blk_bytes(t) != 0; }

static void free_mat(DevMat &m) {
    if (m.qs) cudaFree(m.qs);
    if (m.d) cudaFree(m.d);
    m = DevMat();
}

// Upload a [rows, cols] tensor from host GGUF bytes into the planar device layout.
static int upload_mat(DevMat &m, int type, int64_t rows, int64_t cols, const void *host, size_t nbytes, cudaStream_t st) {
    cudaGetLastError();  // drop a stale (non-sticky) error some other library in the process may have left behind on this thread
    if (!type_supported(type)) return fail(NL_ERR_UNSUPPORTED, "unsupported tensor type %d", type);
    if (rows <= 0 || cols <= 0 || cols % blk_elems(type) != 0) return fail(NL_ERR_INVALID, "bad shape %lldx%lld for type %d", (long long)rows, (long long)cols, type);
    if ((type == NL_F16 && cols % 8) || (type == NL_F32 && cols % 4)) return fail(NL_ERR_INVALID, "cols %lld not a multiple of the 16-byte unit", (long long)cols);
    int64_t expect = tensor_nbytes(type, rows * cols);
    if ((int64_t)nbytes != expect) return fail(NL_ERR_INVALID, "tensor has %zu bytes, expected %lld", nbytes, (long long)expect);
    free_mat(m);
    m.type = type; m.rows = rows; m.cols = cols;
    if (type_planar(type)) {
        int64_t nblocks = rows * cols / 32;
        m.qs_bytes = (size_t)nblocks * (type == NL_Q4_0 ? 16 : 32);
        m.d_bytes = (size_t)nblocks * 2;
        uint8_t *staging = nullptr;
        NL_CUDA(cudaMalloc(&staging, nbytes));
        cudaError_t e = cudaMalloc(&m.qs, m.qs_bytes);
        if (e == cudaSuccess) e = cudaMalloc(&m.d, m.d_bytes);
        if (e != cudaSuccess) { cudaFree(staging); free_mat(m); return fail(NL_ERR_OOM, "cudaMalloc: %s", cudaGetErrorString(e)); }
        e = cudaMemcpyAsync(staging, host, nbytes, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) {
            int threads = 256; int64_t grid = (nblocks + threads - 1) / threads;
            if (type == NL_Q4_0) repack_q4_0_kernel<<<(unsigned)grid, threads, 0, st>>>(staging, (uint4 *)m.qs, m.d, nblocks);
            else repack_q8_0_kernel<<<(unsigned)grid, threads, 0, st>>>(staging, (uint4 *)m.qs, m.d, nblocks);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        cudaFree(staging);
        if (e != cudaSuccess) { free_mat(m); return fail(NL_ERR_CUDA, "upload/repack: %s", cudaGetErrorString(e)); }
    } else {
        m.qs_bytes = nbytes;
        cudaError_t e = cudaMalloc(&m.qs, nbytes);
        if (e != cudaSuccess) { free_mat(m); return fail(NL_ERR_OOM, "cudaMalloc: %s", cudaGetErrorString(e)); }
        e = cudaMemcpyAsync(m.qs, host, nbytes, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { free_mat(m); return fail(NL_ERR_CUDA, "upload: %s", cudaGetErrorString(e)); }
    }
    return NL_OK;
}

// planar/raw device matrix -> fp32 (n = rows*cols elements)
static int dequant_mat(const DevMat &m, float *out, cudaStream_t st) {
    int64_t n = m.rows * m.cols;
    int threads = 256;
    if (m.type == NL_Q4_0) dequant_q4_0_kernel<<<(unsigned)((n / 32 + threads - 1) / threads), threads, 0, st>>>((const uint4 *)m.qs, m.d, out, n / 32);
    else if (m.type == NL_Q8_0) dequant_q8_0_kernel<<<(unsigned)((n / 32 + threads - 1) / threads), threads, 0, st>>>((const uint4 *)m.qs, m.d, out, n / 32);
    else if (m.type == NL_F16) dequant_f16_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, st>>>((const __half *)m.qs, out, n);
    else if (m.type == NL_F32) NL_CUDA(cudaMemcpyAsync(out, m.qs, n * 4, cudaMemcpyDeviceToDevice, st));
    else dequant_raw_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, st>>>(m.type, m.qs, out, n);
    NL_CUDA(cudaGetLastError());
    return NL_OK;
}

// ---- GEMV dispatch on (type, batch) ----
struct MatRef { const DevMat *w; const DevMat *w2; const float *bias; float *out; int out_stride; };

static int gemv_multi(const MatRef *mats, int nmat, const float *x, int x_stride, int batch, int epi, cudaStream_t st) {
    const DevMat &w0 = *mats[0].w;
    const int type = w0.type;
    for (int b0 = 0; b0 < batch;) {
        int nb = (batch - b0 >= 4) ? 4 : (batch - b0 >= 2 ? 2 : 1);
        if (type == NL_Q4_0 || type == NL_Q8_0 || type == NL_F16 || type == NL_F32) {
            GemvArgs a; memset(&a, 0, sizeof a);
            a.nseg = nmat; a.x = x + (int64_t)b0 * x_stride; a.x_stride = x_stride; a.cols = (int)w0.cols;
            for (int i = 0; i < nmat; i++) {
                const DevMat &w = *mats[i].w;
                if (w.type != type || w.cols != w0.cols) return fail(NL_ERR_INVALID, "fused GEMV segments differ in type/cols");
                a.seg[i].qs = w.qs; a.seg[i].d = w.d;
                if (epi == EPI_SWIGLU) { a.seg[i].qs2 = mats[i].w2->qs; a.seg[i].d2 = mats[i].w2->d; }
                a.seg[i].bias = mats[i].bias; a.seg[i].out = mats[i].out + (int64_t)b0 * mats[i].out_stride;
                a.seg[i].rows = (int)w.rows; a.seg[i].out_stride = mats[i].out_stride;
            }
            int rc = type == NL_Q4_0 ? launch_gemv_q4_0(a, nb, epi, st) : type == NL_Q8_0 ? launch_gemv_q8_0(a, nb, epi, st)
                   : type == NL_F16 ? launch_gemv_f16(a, nb, epi, st) : launch_gemv_f32(a, nb, epi, st);
            if (rc) return fail(NL_ERR_INVALID, "gemv launch: bad batch %d", nb);
        } else {
            if (epi == EPI_SWIGLU) return fail(NL_ERR_STATE, "internal: swiglu epilogue on raw-block type");
            for (int i = 0; i < nmat; i++) {
                const DevMat &w = *mats[i].w;
                int warps = 8; unsigned grid = (unsigned)((w.rows + warps - 1) / warps);
                const float *xx = x + (int64_t)b0 * x_stride; float *oo = mats[i].out + (int64_t)b0 * mats[i].out_stride;
                if (nb == 4) gemv_raw_kernel<4><<<grid, warps * 32, 0, st>>>(w.type, w.qs, xx, x_stride, oo, mats[i].out_stride, mats[i].bias, (int)w.rows, (int)w.cols, epi);
                else if (nb == 2) gemv_raw_kernel<2><<<grid, warps * 32, 0, st>>>(w.type, w.qs, xx, x_stride, oo, mats[i].out_stride, mats[i].bias, (int)w.rows, (int)w.cols, epi);
                else gemv_raw_kernel<1><<<grid, warps * 32, 0, st>>>(w.type, w.qs, xx, x_stride, oo, mats[i].out_stride, mats[i].bias, (int)w.rows, (int)w.cols, epi);
            }
        }
        b0 += nb;
    }
    NL_CUDA(cudaGetLastError());
    return NL_OK;
}

__global__ void swiglu_kernel(float *__restrict__ hb, const float *__restrict__ hb2, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) hb[i] = silu_f(hb[i]) * hb2[i];
}


struct GemvOpts { int num_sms; bool stream; bool pdl; };
static GemvOpts default_opts(int device) {
    GemvOpts o{148, true, true};
    cudaDeviceGetAttribute(&o.num_sms, cudaDevAttrMultiProcessorCount, device);
    if (getenv("NL_NO_STREAM")) o.stream = false;   // A/B switch: fall back to the simple per-row-tile GEMV
    if (getenv("NL_NO_PDL")) o.pdl = false;
    return o;
}

// One (possibly multi-segment) matmulDispatch.  norm_w != nullptr asks for out = W · RMSNorm(x; norm_w) (go/quant.go:597-607):
// fused into the streaming kernel's prologue when it applies, otherwise done by rmsnorm_kernel into `scratch`.
// Returns the number of kernel launches issued through *launches.
static int gemv_dispatch(const MatRef *mats, int nmat, const float *x, int x_stride, int batch, int epi, const float *norm_w, float eps,
                         float *scratch, const GemvOpts &o, cudaStream_t st, int *launches) {
    const DevMat &w0 = *mats[0].w;
    const int type = w0.type, cols = (int)w0.cols;
    const int NM = epi == EPI_SWIGLU ? 2 : 1;
    bool same = true;
    for (int i = 0; i < nmat; i++) same = same && mats[i].w->type == type && mats[i].w->cols == w0.cols && (NM == 1 || (mats[i].w2->type == type && mats[i].w2->cols == w0.cols));
    if (o.stream && batch == 1 && same) {
        int rows[3];
        for (int i = 0; i < nmat; i++) rows[i] = (int)mats[i].w->rows;
        StreamPlan p = plan_stream(type, rows, nmat, cols, NM, o.num_sms);
        if (p.ok) {
            StreamArgs a; memset(&a, 0, sizeof a);
            a.nseg = nmat; a.x = x; a.norm_w = norm_w; a.eps = eps; a.cols = cols; a.nb = p.nb; a.nb_pad = p.nb_pad; a.RG = p.RG; a.T = p.T;
            a.stages = p.stages; a.stage_bytes = p.stage_bytes; a.epi = epi == EPI_RESID ? SEPI_RESID : epi == EPI_SWIGLU ? SEPI_SWIGLU : SEPI_STORE;
            int tiles = 0;
            for (int i = 0; i < nmat; i++) {
                const DevMat &w = *mats[i].w;
                a.seg[i].qs = w.qs; a.seg[i].d = w.d;
                if (NM == 2) { a.seg[i].qs2 = mats[i].w2->qs; a.seg[i].d2 = mats[i].w2->d; }
                a.seg[i].bias = mats[i].bias; a.seg[i].out = mats[i].out; a.seg[i].rows = rows[i]; a.seg[i].tile_begin = tiles;
                tiles += (rows[i] + p.T - 1) / p.T;
            }
            a.total_tiles = tiles;
            const int act = norm_w ? ACT_RMSNORM : ACT_NONE;
            int rc = type == NL_Q4_0 ? launch_stream_q4_0(a, NM, p.RPT, act, p.grid, p.smem, st, o.pdl)
                   : type == NL_Q8_0 ? launch_stream_q8_0(a, NM, p.RPT, act, p.grid, p.smem, st, o.pdl)
                                     : launch_stream_f16(a, NM, p.RPT, act, p.grid, p.smem, st, o.pdl);
            if (rc) return fail(NL_ERR_CUDA, "stream gemv launch failed: %s", cudaGetErrorString(cudaGetLastError()));
            if (launches) (*launches)++;
            return NL_OK;
        }
    }
    // ---- fallback path ----
    const float *xin = x;
    if (norm_w) {
        const int nt = cols >= 4096 ? 1024 : cols >= 1024 ? 512 : 256;
        rmsnorm_kernel<<<batch, nt, 0, st>>>(x, norm_w, scratch, cols, eps);
        if (launches) (*launches)++;
        xin = scratch;
    }
    const bool simple_ok = type == NL_Q4_0 || type == NL_Q8_0 || type == NL_F16 || type == NL_F32;
    if (same && (simple_ok || epi != EPI_SWIGLU)) {
        if (simple_ok) { int rc = gemv_multi(mats, nmat, xin, x_stride, batch, epi, st); if (rc) return rc; if (launches) (*launches)++; }
        else for (int i = 0; i < nmat; i++) { int rc = gemv_multi(&mats[i], 1, xin, x_stride, batch, epi, st); if (rc) return rc; if (launches) (*launches)++; }
        return NL_OK;
    }
    // mixed types (or SwiGLU over raw-block types): one launch per matrix, SwiGLU as a separate pass
    if (epi == EPI_SWIGLU) return fail(NL_ERR_STATE, "internal: unfused swiglu must be issued by the caller");
    for (int i = 0; i < nmat; i++) { int rc = gemv_multi(&mats[i], 1, xin, x_stride, batch, epi, st); if (rc) return rc; if (launches) (*launches)++; }
    return NL_OK;
}


// C[T, rows] (=|+=) A[T, cols] · W^T on the tensor cores (nl_gemm.cuh).  a_hi / a_lo: bf16 planes of A, already split.
static bool gemm_v1() { static const bool v = getenv("NL_GEMM_V1") != nullptr; return v; }
#ifndef NL_GEMM_ATILE_DEFAULT
#define NL_GEMM_ATILE_DEFAULT 1
#endif
// NL_GEMM_ATILE: the producers of a GEMM input write its bf16 planes in tile order (plane_index, nl_common.cuh) and the GEMM's issuing
// thread fetches every 128-row x 32-k tile with one bulk copy, instead of 256 threads issuing 16-byte cp.async for it
static int gemm_atile() {
    static const int v = gemm_v1() ? 0 : (getenv("NL_GEMM_ATILE") ? atoi(getenv("NL_GEMM_ATILE")) != 0 : NL_GEMM_ATILE_DEFAULT);
    return v;
}
static bool gemm_eligible(const DevMat &w) {
    // K in whole quant blocks (the first-generation kernel steps by 64: an 8-way shard of big's down projection has 43 blocks per row)
    return (w.type == NL_Q4_0 || w.type == NL_Q8_0 || w.type == NL_F16) && w.cols % (gemm_v1() ? GM_BK : G2_BK) == 0 && w.cols % 8 == 0;
}
// One launch for up to three matrices that share the input (q|k|v, gate|up): nl_gemm2.cuh.  NL_GEMM_V1=1 keeps the first-generation
// kernel (one launch per matrix) for A/B runs.
struct GemmOut { const DevMat *w; const float *bias; float *c; int ldc; int epi; };
static int gemm_run1(const DevMat &w, const __nv_bfloat16 *a_hi, const __nv_bfloat16 *a_lo, int T, const float *bias, float *c, int ldc, int epi, cudaStream_t st) {
    GemmArgs g; memset(&g, 0, sizeof g);
    g.a_hi = a_hi; g.a_lo = a_lo; g.qs = w.qs; g.d = w.d; g.bias = bias; g.c = c; g.T = T; g.N = (int)w.rows; g.K = (int)w.cols; g.ldc = ldc; g.epi = epi;
    g.swap_lbo_sbo = getenv("NL_GEMM_SWAP") ? 1 : 0;
    int rc = w.type == NL_Q4_0 ? launch_gemm_q4_0(g, st) : w.type == NL_Q8_0 ? launch_gemm_q8_0(g, st) : launch_gemm_f16(g, st);
    if (rc) return fail(NL_ERR_CUDA, "tcgen05 GEMM launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    return NL_OK;
}
constexpr size_t G2_SPLIT_BYTES = 16u << 20;
static int gemm_run_multi(const GemmOut *o, int n, const __nv_bfloat16 *a_hi, const __nv_bfloat16 *a_lo, int T, cudaStream_t st, int *launches = nullptr, float *split = nullptr) {
    bool same = n <= G2_MAX_SEG;
    for (int i = 1; i < n; i++) same = same && o[i].w->type == o[0].w->type && o[i].w->cols == o[0].w->cols;
    const bool v1 = gemm_v1();
    if (v1 || !same) {
        for (int i = 0; i < n; i++) {
            int rc = v1 ? gemm_run1(*o[i].w, a_hi, a_lo, T, o[i].bias, o[i].c, o[i].ldc, o[i].epi, st) : gemm_run_multi(&o[i], 1, a_hi, a_lo, T, st, nullptr, split);
            if (rc) return rc;
            if (launches) (*launches)++;
        }
        return NL_OK;
    }
    Gemm2Args g; memset(&g, 0, sizeof g);
    g.a_hi = a_hi; g.a_lo = a_lo; g.nseg = n; g.T = T; g.K = (int)o[0].w->cols; g.a_tiled = gemm_atile();
    int tiles = 0;
    for (int i = 0; i < n; i++) {
        const DevMat &w = *o[i].w;
        tiles += ((int)w.rows + G2_BN - 1) / G2_BN;
        g.seg[i].qs = w.qs; g.seg[i].d = w.d; g.seg[i].bias = o[i].bias; g.seg[i].c = o[i].c; g.seg[i].N = (int)w.rows; g.seg[i].ldc = o[i].ldc; g.seg[i].epi = o[i].epi;
        g.seg[i].tile_end = tiles;
    }
    g.ksplit = 1; g.ksplit_steps = g.K / 32; g.ldp = tiles * G2_BN; g.part = split;
    if (T <= 128 && split && !getenv("NL_GEMM_NOSPLIT")) {
        // a small batch is bound by how fast the weights stream, and one CTA per 256 rows leaves most SMs idle on the narrow matrices:
        // enough K splits for one CTA per SM, at least four K steps each, partial sums within the scratch
        const int nbk = g.K / 32;
        const int want = 148 / tiles, maxs = nbk / 4 < 16 ? nbk / 4 : 16;   // (one wave: a second one pays the CTA's fixed costs again)
        const int cap = (int)(G2_SPLIT_BYTES / ((size_t)T * g.ldp * 4));
        int S = want < maxs ? want : maxs;
        if (S > cap) S = cap;
        if (S > 1) { g.ksplit_steps = (nbk + S - 1) / S; g.ksplit = (nbk + g.ksplit_steps - 1) / g.ksplit_steps; }
    }
    const int type = o[0].w->type;
    int rc = type == NL_Q4_0 ? launch_gemm2_q4_0(g, st) : type == NL_Q8_0 ? launch_gemm2_q8_0(g, st) : launch_gemm2_f16(g, st);
    if (rc) return fail(NL_ERR_CUDA, "tcgen05 GEMM launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    if (launches) *launches += g.ksplit > 1 ? 2 : 1;
    return NL_OK;
}
static int gemm_run(const DevMat &w, const __nv_bfloat16 *a_hi, const __nv_bfloat16 *a_lo, int T, const float *bias, float *c, int ldc, int epi, cudaStream_t st, float *split = nullptr) {
    const GemmOut o{&w, bias, c, ldc, epi};
    return gemm_run_multi(&o, 1, a_hi, a_lo, T, st, nullptr, split);
}
static int split_planes(const float *x, __nv_bfloat16 *hi, __nv_bfloat16 *lo, int64_t n, int K, cudaStream_t st) {
    const int64_t pairs = (n + 1) / 2;
    split_bf16_kernel<<<(unsigned)((pairs + 255) / 256), 256, 0, st>>>(x, hi, lo, n, K, gemm_atile());
    NL_CUDA(cudaGetLastError());
    return NL_OK;
}

}  // namespace nl

using namespace nl;

// =====================================================================================================
struct Layer {
    float *attn_norm = nullptr, *ffn_norm = nullptr, *bq = nullptr, *bk = nullptr, *bv = nullptr, *bo = nullptr;
    DevMat wq, wk, wv, wo, wgate, wup, wdown;
};

struct nl_model {
    nl_config c;
    int dim = 0, hd = 0, kvd = 0, qdim = 0, B = 1;
    cudaStream_t st = nullptr;
    DevMat tok_embd, output;
    float *output_norm = nullptr;
    std::vector<Layer> L;
    float *gamma = nullptr; int32_t *gamma_map = nullptr;
    // state (all [B][...])
    float *x = nullptr, *xb = nullptr, *xb2 = nullptr, *hb = nullptr, *hb2 = nullptr, *q = nullptr, *k = nullptr, *v = nullptr, *logits = nullptr;
    float *kc = nullptr, *vc = nullptr;  // [B][L][S][kvd]
    float *cos_t = nullptr, *sin_t = nullptr;
    int32_t *d_token = nullptr, *d_pos = nullptr, *d_gen = nullptr, *d_gen_count = nullptr, *d_prompt = nullptr, *d_cursor = nullptr;
    int gen_cap = 0, prompt_cap = 0;
    int32_t *h_stage = nullptr;  // pinned: tokens[B], pos[B]
    float *h_logits = nullptr;   // pinned [B][vocab]
    // device-side sampling (nl_sample.cu): sort scratch, repetition window, result; allocated on first use
    uint32_t *sp_keys = nullptr; int32_t *sp_idx = nullptr, *sp_recent = nullptr, *sp_token = nullptr; int32_t *sp_host = nullptr;
    static constexpr int SP_RECENT_CAP = 4096;
    bool finalized = false;
    std::vector<cudaGraphExec_t> g_fwd, g_step;  // index = batch
    int launches_fwd = 0;
    int64_t weight_bytes = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    GemvOpts opts{148, true, true};
    // tensor parallelism (one process per GPU): local shard sizes and the NVLink exchange window
    int tp = 1, rank = 0, nH = 0, nKV = 0, ffn = 0, lvocab = 0;
    float *partial = nullptr, *logits_local = nullptr;
    uint8_t *tp_win = nullptr; TpLayout tp_lay{}; TpPeers tp_peers{}; bool tp_ready = false;
    unsigned int *d_ar_epoch = nullptr, *d_lg_epoch = nullptr;
    // one-pass prefill workspace (allocated on first use, sized for seq_len rows)
    float *pf_x = nullptr, *pf_qkv = nullptr, *pf_g = nullptr, *pf_u = nullptr; __nv_bfloat16 *pf_hi = nullptr, *pf_lo = nullptr; int pf_cap = 0;
    float *pf_split = nullptr;   // split-K partial sums of the small-batch GEMMs (nl_gemm2.cuh), G2_SPLIT_BYTES
    bool pf_time = false; int pf_launches = 0;   // nl_bench_prefill: event after the token copy, kernels launched by the last prefill
    DevMat lm_view;   // this rank's vocab rows of the LM head (a view into output / tok_embd when tied; never freed)
    unsigned int *d_bar = nullptr; float *part_acc = nullptr, *part_ml = nullptr;   // grid-barrier counters, split-attention partials
    unsigned long long *d_trace = nullptr, *d_trace2 = nullptr;
    // tiled tensor-core decode (nl_tile.cuh): fragment-tiled copies of the Q4_0 matrices, batch 1, single GPU
    bool tile_ok = false; int tile_type = NL_Q4_0;
    std::vector<uint8_t *> tile_bufs;   // per layer: qkv, o, gate/up, down; then the LM head
    // shared activation vectors of the tiled path's barrier mode (NL_TILE_POLL=0, tensor parallel): residual stream (tensor parallel:
    // the two exchange residuals), q | k | v, attention output, SwiGLU output
    uint2 *x_sh = nullptr, *qkv_sh = nullptr, *ao_sh = nullptr, *hb_sh = nullptr;
    float *arena = nullptr; size_t arena_bytes = 0;   // polled single-use activation vectors of the tiled path (nl_tile.cu)
    unsigned int *d_epoch = nullptr;    // launch counter behind the flags
    float2 *amax = nullptr; bool amax_valid = false;   // per-CTA argmax pairs of the LM-head phase
    float *qkv_bias = nullptr;
    TilePhase *d_tphases = nullptr; TileArgs targs; int tile_grid = 0;
    // tensor parallel + polled: one arena per token parity inside the window, a descriptor list / TileArgs / graph pair per parity, strictly
    // alternated by the host (every rank launches the same sequence of forwards)
    TpArena tp_arena{}; bool tp_poll = false; TileArgs targs1; int tok_parity = 0; int *d_lg_want = nullptr; unsigned int *d_pf_epoch = nullptr;
    cudaGraphExec_t g_fwd_odd = nullptr, g_step_odd = nullptr;
};

static int set_dev(const nl_model *m) {
    cudaGetLastError();  // see upload_mat
    NL_CUDA(cudaSetDevice(m->c.device));
    return NL_OK;
}

static int upload_vec(float **dst, int type, int64_t n, const void *host, size_t nbytes, cudaStream_t st) {
    DevMat tmp;
    int rc = upload_mat(tmp, type, 1, n, host, nbytes, st);
    if (rc) return rc;
    if (*dst) cudaFree(*dst);
    *dst = nullptr;
    cudaError_t e = cudaMalloc(dst, n * 4);
    if (e != cudaSuccess) { free_mat(tmp); return fail(NL_ERR_OOM, "cudaMalloc: %s", cudaGetErrorString(e)); }
    rc = dequant_mat(tmp, *dst, st);  // getF32Tensor, go/model.go:268-303
    if (rc == NL_OK && cudaStreamSynchronize(st) != cudaSuccess) rc = fail(NL_ERR_CUDA, "dequant vector failed");
    free_mat(tmp);
    return rc;
}

// ---- tiled tensor-core decode path (nl_tile.cuh) ----
static bool tile_eligible(const DevMat &w) {
    return (w.type == NL_Q4_0 || w.type == NL_Q8_0) && w.rows % 16 == 0 && w.cols % 32 == 0 && w.cols <= (int64_t)TL_MAX_NBG * 128 && w.rows * (w.cols / 32) < (1ll << 31) &&
           w.rows < (1ll << 20);   // (band_of: row groups x grid^2 < 2^32)
}
// Builds one tiled matrix from `n` planar matrices of equal cols: concatenated row-wise (interleave = false) or with their
// 16-row groups interleaved (gate/up: group i of mats[0], group i of mats[1], ...).
static int make_tiles(uint8_t **dst, const DevMat *const *mats, int n, bool interleave, cudaStream_t st) {
    const int nb = (int)(mats[0]->cols / 32), nbg = (nb + 3) / 4;
    int64_t n_rg = 0;
    for (int i = 0; i < n; i++) n_rg += mats[i]->rows / 16;
    NL_CUDA(cudaMalloc(dst, (size_t)n_rg * nbg * tile_bytes(mats[0]->type)));
    int off = 0;
    for (int i = 0; i < n; i++) {
        const DevMat &w = *mats[i];
        if (launch_tile_repack(w.type, w.qs, w.d, (int)w.rows, nb, *dst, nbg, interleave ? i : off, interleave ? n : 1, st)) return fail(NL_ERR_CUDA, "tile repack launch failed");
        off += (int)(w.rows / 16);
    }
    return NL_OK;
}
// in / out / resid are fp32 vectors; *_poll: the vector lives in the single-use arena and is polled for the sentinel (nl_tile.cu)
static void tile_gemv_phase(TilePhase &P, const uint8_t *tiles, int n_rg, int cols, int unit_rg, int rows, int epi, const void *x, int in_poll,
                            const float *norm_w, const float *bias, void *out, int out_poll, const void *resid = nullptr, int resid_poll = 0) {
    memset(&P, 0, sizeof P);
    P.kind = PH_GEMV; P.tiles = tiles; P.n_rg = n_rg; P.nbg = (cols / 32 + 3) / 4; P.nbg_magic = tile_magic(P.nbg); P.unit_rg = unit_rg; P.units = n_rg / unit_rg;
    P.cols = cols; P.rows = rows; P.epi = epi; P.x = (const float *)x; P.norm_w = norm_w; P.bias = bias; P.out = (float *)out;
    P.resid = (const float *)resid; P.in_poll = in_poll; P.out_poll = out_poll; P.resid_poll = resid_poll;
}
__global__ void bump_epoch_kernel(unsigned int *epoch) { *epoch += 1; }

// Phase list of the tiled persistent kernel.  Leaves tile_ok = false when the model does not fit (types other than Q4_0, rows not a
// multiple of 16, head_dim != 64, tensor parallel ...): the per-matrix kernels are used then.
static int build_tiled(nl_model *m) {
    const nl_config &c = m->c;
    m->tile_ok = false;
    if (getenv("NL_NO_TILED")) return NL_OK;
    const bool tpar = m->tp > 1;   // tensor parallel: the shards are tiled; o / down exchange their partials inside the kernel
    if (tpar && (getenv("NL_NO_TILED_TP") || !m->tp_ready)) return NL_OK;
    if (m->hd != 64 || c.n_heads / c.n_kv_heads > MG_MAX_GROUP) return NL_OK;
    const DevMat &outw = tpar ? m->lm_view : (m->output.present() ? m->output : m->tok_embd);
    if (!tile_eligible(outw)) return NL_OK;
    const int wtype = outw.type;   // one kernel instantiation per model: every matrix must be of this type
    bool any_bias = false;
    for (const Layer &ly : m->L) {
        for (const DevMat *w : {&ly.wq, &ly.wk, &ly.wv, &ly.wo, &ly.wgate, &ly.wup, &ly.wdown}) if (!tile_eligible(*w) || w->type != wtype) return NL_OK;
        if (ly.bq || ly.bk || ly.bv) any_bias = true;
    }
    const int G = m->opts.num_sms;
    const int dim = m->dim, kvd = m->kvd, qdim = m->qdim, ffn = m->ffn, nqkv = qdim + 2 * kvd;   // this rank's shard sizes
    const int lvocab = tpar ? m->lvocab : c.vocab_size;
    if (G > 256 || 5 * c.n_layers + 1 > 1024) return NL_OK;   // window areas of the tensor-parallel exchange
    cudaStream_t st = m->st;
    {
        struct { uint2 **p; int n; } shv[4] = {{&m->x_sh, dim}, {&m->qkv_sh, nqkv}, {&m->ao_sh, qdim}, {&m->hb_sh, ffn}};
        for (auto &b : shv) { NL_CUDA(cudaMalloc(b.p, (size_t)b.n * 8)); NL_CUDA(cudaMemset(*b.p, 0, (size_t)b.n * 8)); }
        NL_CUDA(cudaMalloc(&m->d_epoch, 4));
        NL_CUDA(cudaMemset(m->d_epoch, 0, 4));
        NL_CUDA(cudaMalloc(&m->amax, (size_t)G * 8));
        NL_CUDA(cudaMemset(m->amax, 0, (size_t)G * 8));
    }
    if (any_bias) {   // q | k | v biases in the same order as the concatenated rows (absent ones are zero)
        NL_CUDA(cudaMalloc(&m->qkv_bias, (size_t)c.n_layers * nqkv * 4));
        NL_CUDA(cudaMemset(m->qkv_bias, 0, (size_t)c.n_layers * nqkv * 4));
        for (int l = 0; l < c.n_layers; l++) {
            float *b = m->qkv_bias + (size_t)l * nqkv;
            const Layer &ly = m->L[l];
            if (ly.bq) NL_CUDA(cudaMemcpyAsync(b, ly.bq, (size_t)qdim * 4, cudaMemcpyDeviceToDevice, st));
            if (ly.bk) NL_CUDA(cudaMemcpyAsync(b + qdim, ly.bk, (size_t)kvd * 4, cudaMemcpyDeviceToDevice, st));
            if (ly.bv) NL_CUDA(cudaMemcpyAsync(b + qdim + kvd, ly.bv, (size_t)kvd * 4, cudaMemcpyDeviceToDevice, st));
        }
    }
    int rc;
    // Polled single-use activation vectors, no grid barrier (nl_tile.cu, "polled activations"): per layer q|k|v, attention output,
    // post-attention residual, SwiGLU output, layer output live in one arena that record_forward fills with the sentinel before every
    // launch.  Tensor parallel: the arena sits in the IPC window (the peers store their partials of the row-split products into it),
    // one per token parity (nl_tp.cuh, TpArena).  NL_TILE_POLL=0: shared vectors behind release / acquire grid barriers.
    const int POLL = !(getenv("NL_TILE_POLL") && atoi(getenv("NL_TILE_POLL")) == 0) ? 1 : 0;
    // producer-side fragments (nl_tile.cuh): every vector a GEMV reads is published by its producer as that GEMV's fragment image,
    // behind the fp32 copy where the residual path still needs one.
    // Measured (profiles/r02_decode_ab.log): the images do not pay on one GPU -- the split moves from 512 consumer threads to the single
    // finishing warp of the producer, both on the critical path -- so they are opt-in there (NL_TILE_IMG=1).  The polled tensor-parallel
    // path always uses them for its two local hand-overs (attention -> o, gate/up -> down).
    const bool tpoll = tpar && POLL;
    const int IMG = POLL && (tpoll || (getenv("NL_TILE_IMG") && atoi(getenv("NL_TILE_IMG")) != 0)) ? 1 : 0;
    // single GPU, per layer (floats): q|k|v, attention output, post-attention residual, SwiGLU output, layer output; then (IMG) the
    // images of the attention output, the post-attention residual, the SwiGLU output and the layer output
    const size_t f_qkv = 0, f_ao = f_qkv + nqkv, f_xres = f_ao + (IMG ? 0 : qdim), f_hb = f_xres + dim, f_xout = f_hb + (IMG ? 0 : ffn);
    const size_t i_ao = f_xout + dim, i_xres = i_ao + (IMG ? tile_img_bytes(qdim) / 4 : 0), i_hb = i_xres + (IMG ? tile_img_bytes(dim) / 4 : 0),
                 i_xout = i_hb + (IMG ? tile_img_bytes(ffn) / 4 : 0), per_layer = i_xout + (IMG ? tile_img_bytes(dim) / 4 : 0);
    if (POLL && !tpar) {
        m->arena_bytes = (size_t)c.n_layers * per_layer * 4;
        NL_CUDA(cudaMalloc(&m->arena, m->arena_bytes));
        NL_CUDA(cudaMemset(m->arena, 0xFF, m->arena_bytes));
    }
    if (tpoll) {
        if (m->tp_arena.total == 0 || m->tp_lay.arena[1] + m->tp_arena.total > m->tp_lay.total) return fail(NL_ERR_STATE, "internal: tensor-parallel arena not laid out");
        NL_CUDA(cudaMemset(m->tp_win + m->tp_lay.arena[0], 0xFF, m->tp_arena.total));
        NL_CUDA(cudaMemset(m->tp_win + m->tp_lay.arena[1], 0xFF, m->tp_arena.total));
    }
    for (int l = 0; l < c.n_layers; l++) {   // the tiled matrices: per layer qkv, o, gate/up, down; then the LM head
        Layer &ly = m->L[l];
        uint8_t *t = nullptr;
        const DevMat *qkv[3] = {&ly.wq, &ly.wk, &ly.wv}, *o[1] = {&ly.wo}, *gu[2] = {&ly.wgate, &ly.wup}, *dn[1] = {&ly.wdown};
        if ((rc = make_tiles(&t, qkv, 3, false, st))) return rc;
        m->tile_bufs.push_back(t);
        if ((rc = make_tiles(&t, o, 1, false, st))) return rc;
        m->tile_bufs.push_back(t);
        if ((rc = make_tiles(&t, gu, 2, true, st))) return rc;
        m->tile_bufs.push_back(t);
        if ((rc = make_tiles(&t, dn, 1, false, st))) return rc;
        m->tile_bufs.push_back(t);
    }
    {
        uint8_t *t = nullptr;
        const DevMat *lm[1] = {&outw};
        if ((rc = make_tiles(&t, lm, 1, false, st))) return rc;
        m->tile_bufs.push_back(t);
    }
    // tensor parallel, barrier mode: exchange e (o-projection of layer l: e = 2l, down-projection: e = 2l + 1) uses parity e & 1 of the
    // exchange area; its consumer adds the ranks' partials to the residual before it (the embedding for e = 0, else xres2[(e - 1) & 1])
    // and leaves the sum in xres2[e & 1]
    float *xres2[2] = {reinterpret_cast<float *>(m->x_sh), reinterpret_cast<float *>(m->x_sh) + dim};   // 2 x dim floats fit the pair buffer
    auto consume_exchange = [&](TilePhase &Q, int e) {
        Q.in_exch = 1; Q.par = e & 1; Q.prev = e == 0 ? m->x : xres2[(e - 1) & 1]; Q.next = xres2[e & 1];
    };
    // one phase list; tensor parallel + polled: one per token parity `tok_par` (the descriptors carry that parity's arena addresses)
    auto make_list = [&](int tok_par) -> std::vector<TilePhase> {
        std::vector<TilePhase> ph;
        TilePhase P;
        void *xres = (void *)m->x;
        const uint8_t *img_last = nullptr;
        const TpArena &ta = m->tp_arena;
        uint8_t *tbase = tpoll ? m->tp_win + m->tp_lay.arena[tok_par] : nullptr;
        for (int l = 0; l < c.n_layers; l++) {
            Layer &ly = m->L[l];
            uint8_t *t_qkv = m->tile_bufs[4 * l], *t_o = m->tile_bufs[4 * l + 1], *t_gu = m->tile_bufs[4 * l + 2], *t_dn = m->tile_bufs[4 * l + 3];
            const float *qb = m->qkv_bias ? m->qkv_bias + (size_t)l * nqkv : nullptr;
            const float *next_norm = l + 1 < c.n_layers ? m->L[l + 1].attn_norm : m->output_norm;
            if (tpoll) {
                // polled exchange: the o / down finishing warps store this rank's partial into slot `rank` of every rank's part_o / part_d;
                // the consumer of an exchange (gate/up, the next layer's qkv, the LM head) polls the residual before it and the tp
                // partials, sums them in rank order (identical on every rank) and its first CTA leaves the new residual in xres / xout
                uint8_t *al = tbase + (size_t)l * ta.per_layer, *ap = al - ta.per_layer;
                const unsigned long long woff = m->tp_lay.arena[tok_par] + (size_t)l * ta.per_layer;
                float *qkv_l = reinterpret_cast<float *>(al + ta.qkv);
                tile_gemv_phase(P, t_qkv, nqkv / 16, dim, 1, nqkv, TEPI_STORE, m->x, 0, ly.attn_norm, qb, qkv_l, 1);
                if (l > 0) { P.in_exch = 1; P.prev = reinterpret_cast<const float *>(ap + ta.xres); P.prev_poll = 1; P.parts = reinterpret_cast<const float *>(ap + ta.part_d);
                             P.next = reinterpret_cast<float *>(ap + ta.xout); }
                ph.push_back(P);
                memset(&P, 0, sizeof P); P.kind = PH_ATTN; P.layer = l; P.x = qkv_l; P.out = nullptr; P.out_img = al + ta.ao_img; ph.push_back(P);
                tile_gemv_phase(P, t_o, dim / 16, qdim, 1, dim, TEPI_STORE, nullptr, 0, nullptr, ly.bo, nullptr, 0);
                P.in_img = al + ta.ao_img; P.exch_out = 1; P.exch_off = woff + ta.part_o;
                ph.push_back(P);
                tile_gemv_phase(P, t_gu, 2 * (ffn / 16), dim, 2, ffn, TEPI_SWIGLU, nullptr, 0, ly.ffn_norm, nullptr, nullptr, 0);
                P.in_exch = 1; P.prev = l == 0 ? m->x : reinterpret_cast<const float *>(ap + ta.xout); P.prev_poll = l > 0; P.parts = reinterpret_cast<const float *>(al + ta.part_o);
                P.next = reinterpret_cast<float *>(al + ta.xres); P.out_img = al + ta.hb_img;
                ph.push_back(P);
                tile_gemv_phase(P, t_dn, dim / 16, ffn, 1, dim, TEPI_STORE, nullptr, 0, nullptr, nullptr, nullptr, 0);
                P.in_img = al + ta.hb_img; P.exch_out = 1; P.exch_off = woff + ta.part_d;
                ph.push_back(P);
                continue;
            }
            // this layer's vectors: shared buffers, or (polled) its own slice of the arena; layer 0 reads the embedding kernel's plain x
            float *a_l = POLL ? m->arena + (size_t)l * per_layer : nullptr;
            void *qkv_l = POLL ? (void *)(a_l + f_qkv) : (void *)m->qkv_sh, *ao_l = POLL ? (IMG ? nullptr : (void *)(a_l + f_ao)) : (void *)m->ao_sh;
            void *hb_l = POLL ? (IMG ? nullptr : (void *)(a_l + f_hb)) : (void *)m->hb_sh;
            const void *x_in = (l == 0 || !POLL) ? (const void *)m->x : (const void *)(a_l - per_layer + f_xout);   // the layer before's output
            if (POLL) xres = a_l + f_xres;            // post-attention residual (output of the o-projection)
            void *x_out = POLL ? (void *)(a_l + f_xout) : xres;
            uint8_t *img_ao = IMG ? reinterpret_cast<uint8_t *>(a_l + i_ao) : nullptr, *img_xres = IMG ? reinterpret_cast<uint8_t *>(a_l + i_xres) : nullptr;
            uint8_t *img_hb = IMG ? reinterpret_cast<uint8_t *>(a_l + i_hb) : nullptr, *img_xout = IMG ? reinterpret_cast<uint8_t *>(a_l + i_xout) : nullptr;
            const uint8_t *img_xin = (IMG && l > 0) ? reinterpret_cast<const uint8_t *>(a_l - per_layer + i_xout) : nullptr;
            tile_gemv_phase(P, t_qkv, nqkv / 16, dim, 1, nqkv, TEPI_STORE, x_in, POLL && l > 0, ly.attn_norm, qb, qkv_l, POLL);
            P.in_img = img_xin;
            if (tpar && l > 0) consume_exchange(P, 2 * l - 1);
            ph.push_back(P);
            memset(&P, 0, sizeof P); P.kind = PH_ATTN; P.layer = l; P.x = (const float *)qkv_l; P.out = (float *)ao_l; P.out_img = img_ao; ph.push_back(P);
            if (tpar) {
                tile_gemv_phase(P, t_o, dim / 16, qdim, 1, dim, TEPI_STORE, ao_l, 0, nullptr, ly.bo, nullptr, 0);
                P.exch_out = 1; P.par = 0; P.cross = 1;
            } else tile_gemv_phase(P, t_o, dim / 16, qdim, 1, dim, TEPI_RESID, ao_l, POLL, nullptr, ly.bo, xres, POLL, x_in, POLL && l > 0);
            P.in_img = img_ao; P.out_img = img_xres; P.out_nw = IMG ? ly.ffn_norm : nullptr;
            ph.push_back(P);
            tile_gemv_phase(P, t_gu, 2 * (ffn / 16), dim, 2, ffn, TEPI_SWIGLU, xres, POLL, ly.ffn_norm, nullptr, hb_l, POLL);
            P.in_img = img_xres; P.out_img = img_hb;
            if (tpar) consume_exchange(P, 2 * l);
            ph.push_back(P);
            if (tpar) {
                tile_gemv_phase(P, t_dn, dim / 16, ffn, 1, dim, TEPI_STORE, hb_l, 0, nullptr, nullptr, nullptr, 0);
                P.exch_out = 1; P.par = 1; P.cross = 1;
            } else tile_gemv_phase(P, t_dn, dim / 16, ffn, 1, dim, TEPI_RESID, hb_l, POLL, nullptr, nullptr, x_out, POLL, xres, POLL);
            P.in_img = img_hb; P.out_img = img_xout; P.out_nw = IMG ? next_norm : nullptr;
            ph.push_back(P);
            if (IMG) img_last = img_xout;
            if (POLL) xres = x_out;   // what the next layer (or the LM head) reads
        }
        uint8_t *t_lm = m->tile_bufs[4 * (size_t)c.n_layers];
        tile_gemv_phase(P, t_lm, lvocab / 16, dim, 1, lvocab, TEPI_STORE, xres, POLL && !tpar, m->output_norm, nullptr, m->logits, 0);
        P.in_img = img_last;
        if (tpoll) {
            uint8_t *ap = tbase + (size_t)(c.n_layers - 1) * ta.per_layer;
            P.in_exch = 1; P.prev = reinterpret_cast<const float *>(ap + ta.xres); P.prev_poll = 1; P.parts = reinterpret_cast<const float *>(ap + ta.part_d);
            P.next = reinterpret_cast<float *>(ap + ta.xout);
        } else if (tpar) { consume_exchange(P, 2 * c.n_layers - 1); P.cross = 1; }   // vocab shard -> every rank's full logits vector
        ph.push_back(P);
        for (size_t i = 1; i < ph.size(); i++) ph[i].wait_cross = ph[i - 1].cross;
        return ph;
    };
    std::vector<TilePhase> ph = make_list(0);
    if (tpoll) { std::vector<TilePhase> ph1 = make_list(1); ph.insert(ph.end(), ph1.begin(), ph1.end()); }
    const size_t n_ph = tpoll ? ph.size() / 2 : ph.size();
    NL_CUDA(cudaStreamSynchronize(st));
    int nsplit = G / m->nKV;
    if (nsplit < 1) nsplit = 1;
    if (nsplit > MG_MAX_SPLIT) nsplit = MG_MAX_SPLIT;
    NL_CUDA(cudaMalloc(&m->d_tphases, ph.size() * sizeof(TilePhase)));
    NL_CUDA(cudaMemcpy(m->d_tphases, ph.data(), ph.size() * sizeof(TilePhase), cudaMemcpyHostToDevice));
    const size_t n_cnt = n_ph + (size_t)c.n_layers * c.n_kv_heads;   // phase barriers, then per-(layer, kv head) split counters
    NL_CUDA(cudaMalloc(&m->d_bar, n_cnt * sizeof(unsigned int)));
    NL_CUDA(cudaMalloc(&m->part_acc, (size_t)m->nH * nsplit * 64 * 8));   // flagged pairs
    NL_CUDA(cudaMalloc(&m->part_ml, (size_t)m->nH * nsplit * 2 * 8));
    NL_CUDA(cudaMemset(m->part_acc, 0, (size_t)m->nH * nsplit * 64 * 8));
    NL_CUDA(cudaMemset(m->part_ml, 0, (size_t)m->nH * nsplit * 2 * 8));
    TileArgs &a = m->targs;
    memset(&a, 0, sizeof a);
    a.phases = m->d_tphases; a.n_phases = (int)n_ph; a.bar = m->d_bar; a.eps = c.rms_norm_eps; a.inflight = tile_inflight(); a.epoch = m->d_epoch; a.poll = POLL;
    a.poll_ns = tile_env_int("NL_TILE_POLL_NS", 0, 0, 2000); a.att_chunk = tile_env_int("NL_ATT_CHUNK", 96, 16, 96); a.att_hpi = tile_env_int("NL_ATT_HPI", 0, 0, 64); a.amax = m->amax; m->amax_valid = true;
    a.dbg = tile_env_int("NL_TILE_DBG", 0, 0, 3);
    a.l2pf = tile_env_int("NL_TILE_L2PF", 4, 0, 64);   // (measured on big: 4 slots ahead +1 %, 12 and more lose: the prefetches compete with the copies)
    // q | k | v is ONE flagged vector: at.q is its base, at.k / at.v only carry element offsets (nl_tile.cu, attn_item_tiled)
    a.at.q = reinterpret_cast<float *>(m->qkv_sh); a.at.k = a.at.q + qdim; a.at.v = a.at.q + qdim + kvd; a.at.kcache = m->kc; a.at.vcache = m->vc; a.at.cos_t = m->cos_t; a.at.sin_t = m->sin_t;
    a.at.pos = m->d_pos; a.at.part_acc = m->part_acc; a.at.part_ml = m->part_ml;
    a.at.n_heads = m->nH; a.at.n_kv_heads = m->nKV; a.at.seq_len = c.seq_len; a.at.qk_norm = c.qk_norm; a.at.conj = c.rope_conjugate;
    if (tpar) {   // barrier counters, exchange area, logits and argmax pairs live in the IPC window every peer can write
        a.tp = m->tp; a.rank = m->rank; a.dim = dim; a.lvocab = lvocab; a.peers = m->tp_peers;
        a.ar_off = m->tp_lay.ar_data; a.bar_off = m->tp_lay.tile_bar; a.lg_off = m->tp_lay.lg_data; a.amax_off = m->tp_lay.tile_amax;
        a.bar = reinterpret_cast<unsigned int *>(m->tp_win + m->tp_lay.tile_bar);
        a.lg_want = m->d_lg_want;
    }
    a.at.nsplit = nsplit; a.at.out = reinterpret_cast<float *>(m->ao_sh); a.at.eps = c.rms_norm_eps;
    a.at.scale = (float)(1.0 / sqrt((double)m->hd));
    if (getenv("NL_TRACE")) {  // latency forensics: per-CTA, per-phase globaltimer stamps of the last token (dumped by nl_bench_decode)
        NL_CUDA(cudaMalloc(&m->d_trace, (size_t)G * n_ph * 8 * sizeof(unsigned long long)));
        NL_CUDA(cudaMemset(m->d_trace, 0, (size_t)G * n_ph * 8 * sizeof(unsigned long long)));
        a.trace = m->d_trace;
        NL_CUDA(cudaMalloc(&m->d_trace2, (size_t)G * n_ph * 16 * sizeof(unsigned long long)));
        NL_CUDA(cudaMemset(m->d_trace2, 0, (size_t)G * n_ph * 16 * sizeof(unsigned long long)));
        a.trace2 = m->d_trace2;
    }
    m->targs.slim = (POLL && !tpar && !IMG) ? 1 : tpoll ? 2 : 0;
    m->tp_poll = tpoll;
    if (tpoll) {   // parity 1: its own descriptor list and argmax-pair area; the host alternates the two (tile_args_for)
        m->targs.amax_off = m->tp_lay.arena[0] + m->tp_arena.amax;
        m->targs1 = m->targs;
        m->targs1.phases = m->d_tphases + n_ph;
        m->targs1.amax_off = m->tp_lay.arena[1] + m->tp_arena.amax;
    }
    m->tile_grid = G; m->tile_type = wtype;
    m->tile_ok = true;
    return NL_OK;
}

static bool batch_gemm_ok(const nl_model *m, int batch);
static int record_forward_batch_gemm(nl_model *m, int batch);

// The launch sequence of one Forward for `batch` sequences (go/model.go:490-620).  Recorded into a CUDA graph.
static int record_forward(nl_model *m, int batch, int par = 0) {
    const nl_config &c = m->c;
    cudaStream_t st = m->st;
    const int dim = m->dim, hd = m->hd, kvd = m->kvd, qdim = m->qdim, S = c.seq_len, ffn = m->ffn;   // shard sizes when tp > 1
    const bool tpar = m->tp > 1;
    if (batch_gemm_ok(m, batch)) return record_forward_batch_gemm(m, batch);
    int launches = 0;
    {   // 1. embedding (+gamma), model.go:500-507
        dim3 grid((dim + 255) / 256, batch);
        embed_kernel<<<grid, 256, 0, st>>>(m->tok_embd, m->d_token, m->gamma, m->gamma_map, m->x, dim);
        launches++;
    }
    if (batch == 1 && m->tile_ok) {
        // everything after the embedding in ONE persistent tensor-core kernel (nl_tile.cuh); its grid-barrier counters start at zero
        // (tensor parallel: the counters live in the IPC window and are never reset)
        // (polled activations: no barriers; every activation vector of the token starts out as sentinels instead)
        // (tensor parallel + polled: this token runs in arena `par`; the OTHER one is refilled for the next token -- no peer writes it
        // before every rank has finished this token, and everything it held was consumed by the token before, nl_tp.cuh)
        if (m->tp_poll) NL_CUDA(cudaMemsetAsync(m->tp_win + m->tp_lay.arena[par ^ 1], 0xFF, m->tp_arena.total, st));
        else if (m->targs.poll) NL_CUDA(cudaMemsetAsync(m->arena, 0xFF, m->arena_bytes, st));
        else if (m->tp == 1) NL_CUDA(cudaMemsetAsync(m->d_bar, 0, ((size_t)m->targs.n_phases + (size_t)m->c.n_layers * m->c.n_kv_heads) * sizeof(unsigned int), st));
        bump_epoch_kernel<<<1, 1, 0, st>>>(m->d_epoch);   // new flags for this token's activation vectors
        launches++;
        if (launch_tiled(m->tile_type, par ? m->targs1 : m->targs, m->tile_grid, st)) return fail(NL_ERR_CUDA, "tiled decode kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        launches++;
        NL_CUDA(cudaGetLastError());
        m->launches_fwd = launches;
        return NL_OK;
    }
    const GemvOpts &o = m->opts;
    const float eps = c.rms_norm_eps;
    for (int l = 0; l < c.n_layers; l++) {
        Layer &ly = m->L[l];
        {   // attention pre-norm + Q, K, V projections (+bias), model.go:517-527 — one launch when the three share a type
            MatRef r[3] = {{&ly.wq, nullptr, ly.bq, m->q, qdim}, {&ly.wk, nullptr, ly.bk, m->k, kvd}, {&ly.wv, nullptr, ly.bv, m->v, kvd}};
            int rc = gemv_dispatch(r, 3, m->x, dim, batch, EPI_STORE, ly.attn_norm, eps, m->xb, o, st, &launches); if (rc) return rc;
        }
        {   // RoPE, QK-norm, KV write, attention: model.go:530-587
            AttnArgs a;
            a.q = m->q; a.k = m->k; a.v = m->v;
            a.kcache = m->kc + (int64_t)l * S * kvd; a.vcache = m->vc + (int64_t)l * S * kvd;
            a.seq_stride = (int64_t)c.n_layers * S * kvd;
            a.cos_t = m->cos_t; a.sin_t = m->sin_t; a.pos = m->d_pos; a.out = m->xb2;
            a.n_heads = m->nH; a.n_kv_heads = m->nKV; a.seq_len = S; a.qk_norm = c.qk_norm; a.conj = c.rope_conjugate;
            a.eps = c.rms_norm_eps; a.scale = (float)(1.0 / sqrt((double)hd));
            dim3 grid(m->nH, batch);
            if (hd == 64) attn_decode_kernel<64><<<grid, 128, S * sizeof(float), st>>>(a);
            else attn_decode_kernel<128><<<grid, 128, S * sizeof(float), st>>>(a);
            launches++;
        }
        {   // output projection + residual, model.go:590-594
            MatRef r = {&ly.wo, nullptr, ly.bo, tpar ? m->partial : m->x, dim};
            int rc = gemv_dispatch(&r, 1, m->xb2, qdim, batch, tpar ? EPI_STORE : EPI_RESID, nullptr, eps, nullptr, o, st, &launches); if (rc) return rc;
            if (tpar) {   // X += sum over ranks of the partial products (row-split WO), one-shot over NVLink peer memory
                tp_allreduce_resid_kernel<<<1, 1024, 0, st>>>(m->partial, m->x, dim, m->tp_peers, m->tp_lay, m->rank, m->tp, m->d_ar_epoch);
                launches++;
            }
        }
        {   // FFN pre-norm + gate/up + SiLU*up, model.go:597-606
            const int gt = ly.wgate.type;
            bool fusable = gt == ly.wup.type && (gt == NL_Q4_0 || gt == NL_Q8_0 || gt == NL_F16 || gt == NL_F32);
            if (fusable) {
                MatRef r = {&ly.wgate, &ly.wup, nullptr, m->hb, ffn};
                int rc = gemv_dispatch(&r, 1, m->x, dim, batch, EPI_SWIGLU, ly.ffn_norm, eps, m->xb, o, st, &launches); if (rc) return rc;
            } else {
                MatRef g = {&ly.wgate, nullptr, nullptr, m->hb, ffn}, u = {&ly.wup, nullptr, nullptr, m->hb2, ffn};
                int rc = gemv_dispatch(&g, 1, m->x, dim, batch, EPI_STORE, ly.ffn_norm, eps, m->xb, o, st, &launches); if (rc) return rc;
                // (the gate dispatch may have fused the norm into its prologue without touching xb: norm again, the same way)
                rc = gemv_dispatch(&u, 1, m->x, dim, batch, EPI_STORE, ly.ffn_norm, eps, m->xb, o, st, &launches); if (rc) return rc;
                int64_t n = (int64_t)batch * ffn;
                swiglu_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(m->hb, m->hb2, n);
                launches++;
            }
        }
        {   // down projection + residual, model.go:609-612
            MatRef r = {&ly.wdown, nullptr, nullptr, tpar ? m->partial : m->x, dim};
            int rc = gemv_dispatch(&r, 1, m->hb, ffn, batch, tpar ? EPI_STORE : EPI_RESID, nullptr, eps, nullptr, o, st, &launches); if (rc) return rc;
            if (tpar) {
                tp_allreduce_resid_kernel<<<1, 1024, 0, st>>>(m->partial, m->x, dim, m->tp_peers, m->tp_lay, m->rank, m->tp, m->d_ar_epoch);
                launches++;
            }
        }
    }
    {   // final norm + LM head, model.go:616-619
        const DevMat &out = tpar ? m->lm_view : (m->output.present() ? m->output : m->tok_embd);
        MatRef r = {&out, nullptr, nullptr, tpar ? m->logits_local : m->logits, tpar ? m->lvocab : c.vocab_size};
        int rc = gemv_dispatch(&r, 1, m->x, dim, batch, EPI_STORE, m->output_norm, eps, m->xb, o, st, &launches); if (rc) return rc;
        if (tpar) {   // vocab-split LM head: every rank drops its shard into every window (m->logits points into this rank's window)
            tp_allgather_logits_kernel<<<1, 1024, 0, st>>>(m->logits_local, m->lvocab, m->tp_peers, m->tp_lay, m->rank, m->tp, m->d_lg_epoch);
            launches++;
        }
    }
    NL_CUDA(cudaGetLastError());
    m->launches_fwd = launches;
    return NL_OK;
}

// Scratch of the GEMM paths (one-pass prefill, batched decode): activations of up to `rows` token rows as fp32 and as bf16 hi | lo planes.
static int ensure_pf(nl_model *m) {
    // sized once for the longest prompt and the largest batch: captured graphs keep these pointers
    const int rows = m->B > m->c.seq_len ? m->B : m->c.seq_len;
    if (m->pf_cap >= rows) return NL_OK;
    if (m->pf_cap) { if (m->tp == 1) cudaFree(m->pf_x); cudaFree(m->pf_qkv); cudaFree(m->pf_g); cudaFree(m->pf_u); cudaFree(m->pf_hi); cudaFree(m->pf_lo); cudaFree(m->pf_split); m->pf_split = nullptr; m->pf_cap = 0; }
    const int dim = m->dim, qdim = m->qdim, kvd = m->kvd, ffn = m->ffn, ld = qdim + 2 * kvd;
    const size_t T = (size_t)rows;
    const size_t Tt = (T + 127) & ~(size_t)127;   // the bf16 planes in whole 128-row tiles (plane_index order reads whole tiles)
    size_t wide = dim > qdim ? dim : qdim; if ((size_t)ffn > wide) wide = ffn;
    if (m->tp > 1) m->pf_x = reinterpret_cast<float *>(m->tp_win + m->tp_lay.pf_x);   // tensor parallel: the residual rows live in the window (nl_tp.cuh)
    else NL_CUDA(cudaMalloc(&m->pf_x, T * dim * 4));
    NL_CUDA(cudaMalloc(&m->pf_qkv, T * ld * 4));
    NL_CUDA(cudaMalloc(&m->pf_g, T * ffn * 4)); NL_CUDA(cudaMalloc(&m->pf_u, T * ffn * 4));
    NL_CUDA(cudaMalloc(&m->pf_hi, Tt * wide * 2)); NL_CUDA(cudaMalloc(&m->pf_lo, Tt * wide * 2));
    NL_CUDA(cudaMalloc(&m->pf_split, G2_SPLIT_BYTES));
    m->pf_cap = rows;
    return NL_OK;
}

#ifndef NL_BATCH_GEMM_MIN_DEFAULT
#define NL_BATCH_GEMM_MIN_DEFAULT 8
#endif
// Batched decode on the tensor cores (NL_BATCH_GEMM_MIN, default 8): B sequences are B token rows of the tcgen05
// GEMMs in their tall orientation (nl_gemm2.cuh: the weights are the M side, the batch the N side, split K on the narrow matrices; the
// weights are read once per step) instead of B accumulators of the CUDA-core GEMV (re-read once per 4 sequences), with the
// per-sequence decode attention in between.  Measured on goldie Q4_0: B = 8 a tie (2.13 ms per step either way), B = 64 2.45 ms
// against 14.35 ms.
static bool batch_gemm_ok(const nl_model *m, int batch) {
    // NL_BATCH_GEMM_MIN = smallest batch that takes this route (0 = never)
    const int bmin = getenv("NL_BATCH_GEMM_MIN") ? atoi(getenv("NL_BATCH_GEMM_MIN")) : (getenv("NL_BATCH_GEMM") ? 16 : NL_BATCH_GEMM_MIN_DEFAULT);
    if (bmin <= 0 || batch < bmin || batch < 2 || m->tp > 1) return false;
    const DevMat &out = m->output.present() ? m->output : m->tok_embd;
    if (!gemm_eligible(out)) return false;
    for (const Layer &ly : m->L)
        for (const DevMat *w : {&ly.wq, &ly.wk, &ly.wv, &ly.wo, &ly.wgate, &ly.wup, &ly.wdown})
            if (!gemm_eligible(*w)) return false;
    return true;
}
static int record_forward_batch_gemm(nl_model *m, int batch) {
    const nl_config &c = m->c;
    cudaStream_t st = m->st;
    const int dim = m->dim, hd = m->hd, kvd = m->kvd, qdim = m->qdim, S = c.seq_len, ffn = m->ffn, B = batch;
    int launches = 0, rc;
    {
        dim3 grid((dim + 255) / 256, B);
        embed_kernel<<<grid, 256, 0, st>>>(m->tok_embd, m->d_token, m->gamma, m->gamma_map, m->x, dim);
        launches++;
    }
    for (int l = 0; l < c.n_layers; l++) {
        Layer &ly = m->L[l];
        rmsnorm_split_kernel<<<B, 256, 0, st>>>(m->x, ly.attn_norm, m->pf_hi, m->pf_lo, dim, c.rms_norm_eps, gemm_atile());
        {
            const GemmOut o[3] = {{&ly.wq, ly.bq, m->q, qdim, GEPI_STORE}, {&ly.wk, ly.bk, m->k, kvd, GEPI_STORE}, {&ly.wv, ly.bv, m->v, kvd, GEPI_STORE}};
            if ((rc = gemm_run_multi(o, 3, m->pf_hi, m->pf_lo, B, st, nullptr, m->pf_split))) return rc;
        }
        {   // RoPE, QK-norm, KV write, attention per sequence: model.go:530-587
            AttnArgs a;
            a.q = m->q; a.k = m->k; a.v = m->v;
            a.kcache = m->kc + (int64_t)l * S * kvd; a.vcache = m->vc + (int64_t)l * S * kvd;
            a.seq_stride = (int64_t)c.n_layers * S * kvd;
            a.cos_t = m->cos_t; a.sin_t = m->sin_t; a.pos = m->d_pos; a.out = m->xb2;
            a.n_heads = m->nH; a.n_kv_heads = m->nKV; a.seq_len = S; a.qk_norm = c.qk_norm; a.conj = c.rope_conjugate;
            a.eps = c.rms_norm_eps; a.scale = (float)(1.0 / sqrt((double)hd));
            dim3 grid(m->nH, B);
            if (hd == 64) attn_decode_kernel<64><<<grid, 128, S * sizeof(float), st>>>(a);
            else attn_decode_kernel<128><<<grid, 128, S * sizeof(float), st>>>(a);
        }
        if ((rc = split_planes(m->xb2, m->pf_hi, m->pf_lo, (int64_t)B * qdim, qdim, st))) return rc;
        if ((rc = gemm_run(ly.wo, m->pf_hi, m->pf_lo, B, ly.bo, m->x, dim, GEPI_RESID, st, m->pf_split))) return rc;
        rmsnorm_split_kernel<<<B, 256, 0, st>>>(m->x, ly.ffn_norm, m->pf_hi, m->pf_lo, dim, c.rms_norm_eps, gemm_atile());
        {
            const GemmOut o[2] = {{&ly.wgate, nullptr, m->pf_g, ffn, GEPI_STORE}, {&ly.wup, nullptr, m->pf_u, ffn, GEPI_STORE}};
            if ((rc = gemm_run_multi(o, 2, m->pf_hi, m->pf_lo, B, st, nullptr, m->pf_split))) return rc;
        }
        {
            const int64_t ne = (int64_t)B * ffn;
            swiglu_split_kernel<<<(unsigned)((ne / 2 + 255) / 256), 256, 0, st>>>(m->pf_g, m->pf_u, m->pf_hi, m->pf_lo, ne, ffn, gemm_atile());
        }
        if ((rc = gemm_run(ly.wdown, m->pf_hi, m->pf_lo, B, nullptr, m->x, dim, GEPI_RESID, st, m->pf_split))) return rc;
        launches += 12;
    }
    {   // final norm + LM head for every sequence, model.go:616-619
        const DevMat &out = m->output.present() ? m->output : m->tok_embd;
        rmsnorm_split_kernel<<<B, 256, 0, st>>>(m->x, m->output_norm, m->pf_hi, m->pf_lo, dim, c.rms_norm_eps, gemm_atile());
        if ((rc = gemm_run(out, m->pf_hi, m->pf_lo, B, nullptr, m->logits, c.vocab_size, GEPI_STORE, st, m->pf_split))) return rc;
        launches += 2;
    }
    NL_CUDA(cudaGetLastError());
    m->launches_fwd = launches;
    return NL_OK;
}

static int record_advance(nl_model *m, int batch, int par = 0) {
    StepState s{m->d_token, m->d_pos, m->d_gen, m->d_gen_count, m->gen_cap};
    // batch 1 on the tiled path: the LM-head phase of the previous forward left one (max, index) pair per CTA
    // (tensor parallel + polled: the previous token ran in the other parity's arena)
    if (batch == 1 && m->tile_ok && m->amax_valid) {
        const float2 *pairs = m->tp_poll ? reinterpret_cast<const float2 *>(m->tp_win + m->tp_lay.arena[par ^ 1] + m->tp_arena.amax)
                            : m->tp > 1 ? reinterpret_cast<const float2 *>(m->tp_win + m->tp_lay.tile_amax) : m->amax;
        argmax_pairs_advance_kernel<<<1, 32, 0, m->st>>>(pairs, m->tile_grid * m->tp, s);
    }
    else argmax_advance_kernel<<<batch, 1024, 0, m->st>>>(m->logits, m->c.vocab_size, s);
    NL_CUDA(cudaGetLastError());
    return NL_OK;
}

static int build_graphs(nl_model *m, int batch) {
    if ((int)m->g_fwd.size() <= batch) { m->g_fwd.resize(batch + 1, nullptr); m->g_step.resize(batch + 1, nullptr); }
    if (m->g_fwd[batch]) return NL_OK;
    cudaGraph_t g;
    if (batch_gemm_ok(m, batch)) { int rc0 = ensure_pf(m); if (rc0) return rc0; }   // no allocation inside a capture
    NL_CUDA(cudaStreamBeginCapture(m->st, cudaStreamCaptureModeThreadLocal));
    int rc = record_forward(m, batch);
    cudaError_t e = cudaStreamEndCapture(m->st, &g);
    if (rc) { if (e == cudaSuccess) cudaGraphDestroy(g); return rc; }
    NL_CUDA(e);
    NL_CUDA(cudaGraphInstantiate(&m->g_fwd[batch], g, 0));
    cudaGraphDestroy(g);
    // greedy step: argmax/advance, then forward (go/main.go:173-218)
    NL_CUDA(cudaStreamBeginCapture(m->st, cudaStreamCaptureModeThreadLocal));
    rc = record_advance(m, batch);
    if (!rc) rc = record_forward(m, batch);
    e = cudaStreamEndCapture(m->st, &g);
    if (rc) { if (e == cudaSuccess) cudaGraphDestroy(g); return rc; }
    NL_CUDA(e);
    NL_CUDA(cudaGraphInstantiate(&m->g_step[batch], g, 0));
    cudaGraphDestroy(g);
    if (m->tp_poll && batch == 1) {   // the odd-parity twins
        NL_CUDA(cudaStreamBeginCapture(m->st, cudaStreamCaptureModeThreadLocal));
        rc = record_forward(m, 1, 1);
        e = cudaStreamEndCapture(m->st, &g);
        if (rc) { if (e == cudaSuccess) cudaGraphDestroy(g); return rc; }
        NL_CUDA(e);
        NL_CUDA(cudaGraphInstantiate(&m->g_fwd_odd, g, 0));
        cudaGraphDestroy(g);
        NL_CUDA(cudaStreamBeginCapture(m->st, cudaStreamCaptureModeThreadLocal));
        rc = record_advance(m, 1, 1);
        if (!rc) rc = record_forward(m, 1, 1);
        e = cudaStreamEndCapture(m->st, &g);
        if (rc) { if (e == cudaSuccess) cudaGraphDestroy(g); return rc; }
        NL_CUDA(e);
        NL_CUDA(cudaGraphInstantiate(&m->g_step_odd, g, 0));
        cudaGraphDestroy(g);
    }
    return NL_OK;
}
// The graph of the next batch-1 forward (step = argmax / advance first).  Tensor parallel + polled: the two parities alternate.
static cudaGraphExec_t next_graph(nl_model *m, bool step) {
    if (!m->tp_poll) return step ? m->g_step[1] : m->g_fwd[1];
    const int par = m->tok_parity;
    m->tok_parity ^= 1;
    return step ? (par ? m->g_step_odd : m->g_step[1]) : (par ? m->g_fwd_odd : m->g_fwd[1]);
}
// tensor parallel: does the next forward's LM head spread its logits shard over the peers' windows (the caller reads logits), or only
// the argmax pairs (greedy / bench loops)?
static int set_lg_want(nl_model *m, int want) {
    if (m->d_lg_want) NL_CUDA(cudaMemsetAsync(m->d_lg_want, want ? 1 : 0, 4, m->st));
    return NL_OK;
}

// =====================================================================================================
extern "C" {

const char *nl_last_error(void) { return g_err; }
int nl_abi_version(void) { return NL_ABI_VERSION; }

int nl_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    int ok = 0;
    for (int i = 0; i < n; i++) {
        int major = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, i) == cudaSuccess && major == 10) ok++;
    }
    return ok;
}

static int check_device(int dev) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) { cudaGetLastError(); return fail(NL_ERR_CUDA, "no CUDA device available (%s); libnanollama_cuda has no CPU fallback", e == cudaSuccess ? "count=0" : cudaGetErrorString(e)); }
    if (dev < 0 || dev >= n) return fail(NL_ERR_INVALID, "device %d out of range (have %d)", dev, n);
    int major = 0;
    NL_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    if (major != 10) return fail(NL_ERR_CUDA, "device %d has compute capability %d.x; this library is built for sm_100a only", dev, major);
    NL_CUDA(cudaSetDevice(dev));
    return NL_OK;
}

int nl_create(const nl_config *cfg, nl_model **out) {
    if (!cfg || !out) return fail(NL_ERR_INVALID, "null argument");
    *out = nullptr;
    nl_config c = *cfg;
    if (c.head_dim == 0 && c.n_heads > 0) c.head_dim = c.embed_dim / c.n_heads;  // go/model.go:140-142
    if (c.seq_len > 2048) c.seq_len = 2048;                                        // go/model.go:145-148
    if (c.n_kv_heads == 0) c.n_kv_heads = c.n_heads;                               // go/gguf.go:456-458
    if (c.max_batch <= 0) c.max_batch = 1;
    if (c.tp_size <= 0) c.tp_size = 1;
    if (c.n_layers <= 0 || c.embed_dim <= 0 || c.n_heads <= 0 || c.n_kv_heads <= 0 || c.vocab_size <= 0 || c.seq_len <= 0 || c.interm_size <= 0)
        return fail(NL_ERR_INVALID, "config has a non-positive dimension (layers=%d dim=%d heads=%d kv=%d vocab=%d seq=%d ffn=%d)",
                    c.n_layers, c.embed_dim, c.n_heads, c.n_kv_heads, c.vocab_size, c.seq_len, c.interm_size);
    if (c.n_heads % c.n_kv_heads) return fail(NL_ERR_INVALID, "n_heads %d not a multiple of n_kv_heads %d", c.n_heads, c.n_kv_heads);
    if (c.head_dim != 64 && c.head_dim != 128) return fail(NL_ERR_UNSUPPORTED, "head_dim %d (supported: 64, 128)", c.head_dim);
    if (c.embed_dim % 32 || c.interm_size % 32) return fail(NL_ERR_INVALID, "embed_dim/interm_size must be multiples of 32");
    if (c.tp_size != 1 && c.tp_size != 2 && c.tp_size != 4 && c.tp_size != 8) return fail(NL_ERR_INVALID, "tp_size %d (supported: 1, 2, 4, 8)", c.tp_size);
    if (c.tp_rank < 0 || c.tp_rank >= c.tp_size) return fail(NL_ERR_INVALID, "tp_rank %d out of range [0,%d)", c.tp_rank, c.tp_size);
    if (c.tp_size > 1) {
        const int n = c.tp_size;
        // column-split q/k/v/gate/up by heads / ffn rows, row-split o/down along whole 32-element blocks, vocab-split LM head
        if (c.n_heads % n || c.n_kv_heads % n) return fail(NL_ERR_INVALID, "tp_size %d does not divide n_heads %d / n_kv_heads %d evenly", n, c.n_heads, c.n_kv_heads);
        if (c.interm_size % (32 * n) || (c.n_heads / n * c.head_dim) % 32) return fail(NL_ERR_INVALID, "tp_size %d: shard boundaries would cut a 32-element quant block", n);
        if (c.vocab_size % (4 * n)) return fail(NL_ERR_INVALID, "tp_size %d does not divide vocab_size %d into float4-aligned shards", n, c.vocab_size);
        if (c.max_batch != 1) return fail(NL_ERR_UNSUPPORTED, "tensor parallelism is built for batch 1");
    }
    if (c.max_batch > 64) return fail(NL_ERR_INVALID, "max_batch %d > 64", c.max_batch);
    int rc = check_device(c.device);
    if (rc) return rc;
    nl_model *m = new nl_model();
    m->c = c; m->dim = c.embed_dim; m->hd = c.head_dim; m->kvd = c.n_kv_heads * c.head_dim; m->qdim = c.n_heads * c.head_dim; m->B = c.max_batch;
    m->tp = c.tp_size; m->rank = c.tp_rank; m->nH = c.n_heads / m->tp; m->nKV = c.n_kv_heads / m->tp; m->ffn = c.interm_size / m->tp; m->lvocab = c.vocab_size / m->tp;
    m->qdim = m->nH * c.head_dim; m->kvd = m->nKV * c.head_dim;   // from here on qdim / kvd / ffn are this rank's shard sizes
    m->L.resize(c.n_layers);
    m->opts = default_opts(c.device);
    cudaError_t e = cudaStreamCreateWithFlags(&m->st, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete m; return fail(NL_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)); }
    *out = m;
    return NL_OK;
}

int nl_get_config(const nl_model *m, nl_config *out) {
    if (!m || !out) return fail(NL_ERR_INVALID, "null argument");
    *out = m->c;
    return NL_OK;
}

int nl_upload_tensor(nl_model *m, int slot, int layer, uint32_t type, int64_t rows, int64_t cols, const void *host, size_t nbytes) {
    if (!m || !host) return fail(NL_ERR_INVALID, "null argument");
    if (m->finalized) return fail(NL_ERR_STATE, "upload after nl_finalize");
    int rc = set_dev(m); if (rc) return rc;
    const nl_config &c = m->c;
    // shapes of the FULL tensors as they sit in the GGUF; with tp_size > 1 only this rank's shard is kept
    const int dim = m->dim, kvd = c.n_kv_heads * c.head_dim, qdim = c.n_heads * c.head_dim, ffn = c.interm_size;
    const int tp = m->tp, rk = m->rank;
    if (slot >= NL_ATTN_NORM && (layer < 0 || layer >= c.n_layers)) return fail(NL_ERR_INVALID, "layer %d out of range", layer);
    Layer *ly = slot >= NL_ATTN_NORM ? &m->L[layer] : nullptr;
    DevMat *mat = nullptr; float **vec = nullptr; int64_t er = 1, ec = 0;
    enum { FULL, ROWS, COLS } split = FULL;
    switch (slot) {
    case NL_TOK_EMBD: mat = &m->tok_embd; er = c.vocab_size; ec = dim; break;             // any token may be looked up: kept whole
    case NL_OUTPUT: mat = &m->output; er = c.vocab_size; ec = dim; split = ROWS; break;
    case NL_OUTPUT_NORM: vec = &m->output_norm; ec = dim; break;
    case NL_ATTN_NORM: vec = &ly->attn_norm; ec = dim; break;
    case NL_FFN_NORM: vec = &ly->ffn_norm; ec = dim; break;
    case NL_WQ: mat = &ly->wq; er = qdim; ec = dim; split = ROWS; break;
    case NL_WK: mat = &ly->wk; er = kvd; ec = dim; split = ROWS; break;
    case NL_WV: mat = &ly->wv; er = kvd; ec = dim; split = ROWS; break;
    case NL_WO: mat = &ly->wo; er = dim; ec = qdim; split = COLS; break;
    case NL_WGATE: mat = &ly->wgate; er = ffn; ec = dim; split = ROWS; break;
    case NL_WUP: mat = &ly->wup; er = ffn; ec = dim; split = ROWS; break;
    case NL_WDOWN: mat = &ly->wdown; er = dim; ec = ffn; split = COLS; break;
    case NL_BQ: vec = &ly->bq; ec = qdim; split = ROWS; break;
    case NL_BK: vec = &ly->bk; ec = kvd; split = ROWS; break;
    case NL_BV: vec = &ly->bv; ec = kvd; split = ROWS; break;
    case NL_BO: vec = &ly->bo; ec = dim; break;
    default: return fail(NL_ERR_INVALID, "unknown tensor slot %d", slot);
    }
    if (tp == 1) split = FULL;
    if (vec) {
        if (rows * cols != ec) return fail(NL_ERR_INVALID, "slot %d: %lld elements, expected %lld", slot, (long long)(rows * cols), (long long)ec);
        if (slot == NL_BO && rk != 0) return NL_OK;   // the output-projection bias is added once, by rank 0
        rc = upload_vec(vec, (int)type, ec, host, nbytes, m->st);
        if (rc || split == FULL) return rc;
        const int64_t ln = ec / tp;                    // keep elements [rk*ln, (rk+1)*ln)
        float *shard = nullptr;
        NL_CUDA(cudaMalloc(&shard, ln * 4));
        NL_CUDA(cudaMemcpy(shard, *vec + rk * ln, ln * 4, cudaMemcpyDeviceToDevice));
        cudaFree(*vec); *vec = shard;
        return NL_OK;
    }
    if (rows != er || cols != ec) return fail(NL_ERR_INVALID, "slot %d: shape %lldx%lld, expected %lldx%lld", slot, (long long)rows, (long long)cols, (long long)er, (long long)ec);
    if (!type_supported((int)type)) return fail(NL_ERR_UNSUPPORTED, "unsupported tensor type %u", type);
    if ((int64_t)nbytes != tensor_nbytes((int)type, rows * cols)) return fail(NL_ERR_INVALID, "tensor has %zu bytes, expected %lld", nbytes, (long long)tensor_nbytes((int)type, rows * cols));
    const int64_t row_bytes = cols / blk_elems((int)type) * blk_bytes((int)type);
    if (split == ROWS) {
        const int64_t lr = rows / tp;
        return upload_mat(*mat, (int)type, lr, cols, (const uint8_t *)host + rk * lr * row_bytes, (size_t)(lr * row_bytes), m->st);
    }
    if (split == COLS) {   // whole quant blocks [rk*lc/be, (rk+1)*lc/be) of every row, gathered on the host
        const int64_t lc = cols / tp;
        if (lc % blk_elems((int)type)) return fail(NL_ERR_INVALID, "column shard %lld cuts a quant block of type %u", (long long)lc, type);
        const int64_t lb = lc / blk_elems((int)type) * blk_bytes((int)type);
        std::vector<uint8_t> tmp((size_t)(rows * lb));
        for (int64_t r = 0; r < rows; r++) memcpy(tmp.data() + r * lb, (const uint8_t *)host + r * row_bytes + rk * lb, (size_t)lb);
        return upload_mat(*mat, (int)type, rows, lc, tmp.data(), tmp.size(), m->st);
    }
    return upload_mat(*mat, (int)type, rows, cols, host, nbytes, m->st);
}

int nl_set_gamma(nl_model *m, const float *rows, int32_t n_rows, const int32_t *token_to_row) {
    if (!m) return fail(NL_ERR_INVALID, "null argument");
    int rc = set_dev(m); if (rc) return rc;
    // graphs bake the gamma pointers in: drop them so the next forward re-captures
    NL_CUDA(cudaStreamSynchronize(m->st));
    for (auto &g : m->g_fwd) if (g) { cudaGraphExecDestroy(g); g = nullptr; }
    for (auto &g : m->g_step) if (g) { cudaGraphExecDestroy(g); g = nullptr; }
    if (m->g_fwd_odd) { cudaGraphExecDestroy(m->g_fwd_odd); m->g_fwd_odd = nullptr; }
    if (m->g_step_odd) { cudaGraphExecDestroy(m->g_step_odd); m->g_step_odd = nullptr; }
    if (m->gamma) { cudaFree(m->gamma); m->gamma = nullptr; }
    if (m->gamma_map) { cudaFree(m->gamma_map); m->gamma_map = nullptr; }
    if (rows) {
        if (n_rows <= 0 || !token_to_row) return fail(NL_ERR_INVALID, "gamma: bad arguments");
        for (int t = 0; t < m->c.vocab_size; t++)   // the embedding kernel indexes the rows with these: a public ABI checks them
            if (token_to_row[t] < -1 || token_to_row[t] >= n_rows) return fail(NL_ERR_INVALID, "gamma: token_to_row[%d] = %d outside [-1, %d)", t, token_to_row[t], n_rows);
        NL_CUDA(cudaMalloc(&m->gamma, (size_t)n_rows * m->dim * 4));
        NL_CUDA(cudaMalloc(&m->gamma_map, (size_t)m->c.vocab_size * 4));
        NL_CUDA(cudaMemcpy(m->gamma, rows, (size_t)n_rows * m->dim * 4, cudaMemcpyHostToDevice));
        NL_CUDA(cudaMemcpy(m->gamma_map, token_to_row, (size_t)m->c.vocab_size * 4, cudaMemcpyHostToDevice));
    }
    // nl_generate_greedy / nl_prefill / nl_bench_decode launch the batch-1 graphs directly: re-capture them now
    if (m->finalized && (m->tp == 1 || m->tp_ready)) return build_graphs(m, 1);
    return NL_OK;
}

int nl_finalize(nl_model *m) {
    if (!m) return fail(NL_ERR_INVALID, "null argument");
    if (m->finalized) return NL_OK;
    int rc = set_dev(m); if (rc) return rc;
    const nl_config &c = m->c;
    // every tensor loadWeights requires (go/model.go:177-265)
    if (!m->tok_embd.present()) return fail(NL_ERR_STATE, "token_embd.weight: tensor not uploaded");
    if (!m->output_norm) return fail(NL_ERR_STATE, "output_norm.weight: tensor not uploaded");
    for (int l = 0; l < c.n_layers; l++) {
        Layer &ly = m->L[l];
        const char *miss = !ly.attn_norm ? "attn_norm" : !ly.ffn_norm ? "ffn_norm" : !ly.wq.present() ? "attn_q" : !ly.wk.present() ? "attn_k"
                         : !ly.wv.present() ? "attn_v" : !ly.wo.present() ? "attn_output" : !ly.wgate.present() ? "ffn_gate"
                         : !ly.wup.present() ? "ffn_up" : !ly.wdown.present() ? "ffn_down" : nullptr;
        if (miss) return fail(NL_ERR_STATE, "layer %d %s: tensor not uploaded", l, miss);
    }
    const int B = m->B, dim = m->dim, kvd = m->kvd, qdim = m->qdim, S = c.seq_len, half = m->hd / 2, ffn = m->ffn;
    auto alloc = [&](float **p, size_t n) -> int { NL_CUDA(cudaMalloc(p, n * 4)); NL_CUDA(cudaMemset(*p, 0, n * 4)); return NL_OK; };
    if ((rc = alloc(&m->x, (size_t)B * dim)) || (rc = alloc(&m->xb, (size_t)B * dim)) || (rc = alloc(&m->xb2, (size_t)B * qdim)) ||
        (rc = alloc(&m->hb, (size_t)B * ffn)) || (rc = alloc(&m->hb2, (size_t)B * ffn)) || (rc = alloc(&m->q, (size_t)B * qdim)) ||
        (rc = alloc(&m->k, (size_t)B * kvd)) || (rc = alloc(&m->v, (size_t)B * kvd)) || (m->tp == 1 && (rc = alloc(&m->logits, (size_t)B * c.vocab_size))) ||
        (rc = alloc(&m->kc, (size_t)B * c.n_layers * S * kvd)) || (rc = alloc(&m->vc, (size_t)B * c.n_layers * S * kvd)))
        return rc;
    // RoPE tables: float64 pow/cos/sin rounded to fp32 (go/model.go:346-358)
    std::vector<float> hc((size_t)S * half), hs((size_t)S * half);
    const double theta = (double)c.rope_theta;
    for (int pos = 0; pos < S; pos++)
        for (int i = 0; i < half; i++) {
            double freq = 1.0 / pow(theta, (double)(2 * i) / (double)m->hd);
            double angle = (double)pos * freq;
            hc[(size_t)pos * half + i] = (float)cos(angle);
            hs[(size_t)pos * half + i] = (float)sin(angle);
        }
    NL_CUDA(cudaMalloc(&m->cos_t, hc.size() * 4)); NL_CUDA(cudaMalloc(&m->sin_t, hs.size() * 4));
    NL_CUDA(cudaMemcpy(m->cos_t, hc.data(), hc.size() * 4, cudaMemcpyHostToDevice));
    NL_CUDA(cudaMemcpy(m->sin_t, hs.data(), hs.size() * 4, cudaMemcpyHostToDevice));
    m->gen_cap = S + 8; m->prompt_cap = S + 8;
    NL_CUDA(cudaMalloc(&m->d_token, B * 4)); NL_CUDA(cudaMalloc(&m->d_pos, B * 4)); NL_CUDA(cudaMalloc(&m->d_gen_count, B * 4));
    NL_CUDA(cudaMalloc(&m->d_gen, (size_t)B * m->gen_cap * 4)); NL_CUDA(cudaMalloc(&m->d_prompt, (size_t)m->prompt_cap * 4)); NL_CUDA(cudaMalloc(&m->d_cursor, 4));
    NL_CUDA(cudaMemset(m->d_token, 0, B * 4)); NL_CUDA(cudaMemset(m->d_pos, 0, B * 4)); NL_CUDA(cudaMemset(m->d_gen_count, 0, B * 4)); NL_CUDA(cudaMemset(m->d_cursor, 0, 4));
    NL_CUDA(cudaMallocHost(&m->h_stage, 2 * 64 * 4));
    NL_CUDA(cudaMallocHost(&m->h_logits, (size_t)B * c.vocab_size * 4));
    NL_CUDA(cudaEventCreate(&m->ev0)); NL_CUDA(cudaEventCreate(&m->ev1));
    int64_t wb = m->tok_embd.bytes() + m->output.bytes() + (int64_t)dim * 4;
    for (auto &ly : m->L) wb += ly.wq.bytes() + ly.wk.bytes() + ly.wv.bytes() + ly.wo.bytes() + ly.wgate.bytes() + ly.wup.bytes() + ly.wdown.bytes() + 2 * (int64_t)dim * 4;
    m->weight_bytes = wb;
    if (S * sizeof(float) > 48 * 1024) return fail(NL_ERR_INVALID, "seq_len too large for the attention kernel");
    if (m->tp > 1) {
        // exchange window (exported to the peers with CUDA IPC), partial-product buffer, local logits shard
        m->tp_arena = tp_arena(m->tp, dim, qdim + 2 * kvd, tile_img_bytes(qdim), tile_img_bytes(ffn), c.n_layers);
        m->tp_lay = tp_layout(m->tp, dim, c.vocab_size, m->tp_arena.total, c.seq_len);
        NL_CUDA(cudaMalloc(&m->tp_win, m->tp_lay.total));
        NL_CUDA(cudaMemset(m->tp_win, 0, m->tp_lay.total));
        NL_CUDA(cudaMalloc(&m->d_lg_want, 4));
        NL_CUDA(cudaMemset(m->d_lg_want, 1, 4));
        NL_CUDA(cudaMalloc(&m->d_pf_epoch, 4));
        NL_CUDA(cudaMemset(m->d_pf_epoch, 0, 4));
        m->logits = reinterpret_cast<float *>(m->tp_win + m->tp_lay.lg_data);
        if ((rc = alloc(&m->partial, (size_t)dim)) || (rc = alloc(&m->logits_local, (size_t)m->lvocab))) return rc;
        NL_CUDA(cudaMalloc(&m->d_ar_epoch, 4)); NL_CUDA(cudaMalloc(&m->d_lg_epoch, 4));
        NL_CUDA(cudaMemset(m->d_ar_epoch, 0, 4)); NL_CUDA(cudaMemset(m->d_lg_epoch, 0, 4));
        // LM head shard: uploaded output.weight is already this rank's rows; a tied embedding is viewed at its vocab slice
        if (m->output.present()) m->lm_view = m->output;
        else {
            const DevMat &e = m->tok_embd;
            m->lm_view = e; m->lm_view.rows = m->lvocab;
            if (e.type == NL_Q4_0 || e.type == NL_Q8_0) {
                const int64_t nb = (int64_t)m->rank * m->lvocab * (e.cols / 32);
                m->lm_view.qs = e.qs + nb * (e.type == NL_Q4_0 ? 16 : 32); m->lm_view.d = e.d + nb;
            } else {
                m->lm_view.qs = e.qs + (int64_t)m->rank * m->lvocab * (e.cols / blk_elems(e.type)) * blk_bytes(e.type);
            }
        }
        m->finalized = true;   // graphs are captured by nl_tp_import_handles once the peers' windows are mapped
        return NL_OK;
    }
    rc = build_tiled(m);
    if (rc) return rc;
    rc = build_graphs(m, 1);
    if (rc) return rc;
    m->finalized = true;
    return NL_OK;
}

void nl_destroy(nl_model *m) {
    if (!m) return;
    cudaSetDevice(m->c.device);
    if (m->st) cudaStreamSynchronize(m->st);
    for (auto g : m->g_fwd) if (g) cudaGraphExecDestroy(g);
    for (auto g : m->g_step) if (g) cudaGraphExecDestroy(g);
    if (m->g_fwd_odd) cudaGraphExecDestroy(m->g_fwd_odd);
    if (m->g_step_odd) cudaGraphExecDestroy(m->g_step_odd);
    if (m->d_lg_want) cudaFree(m->d_lg_want);
    if (m->d_pf_epoch) cudaFree(m->d_pf_epoch);
    free_mat(m->tok_embd); free_mat(m->output);
    for (auto &ly : m->L) {
        free_mat(ly.wq); free_mat(ly.wk); free_mat(ly.wv); free_mat(ly.wo); free_mat(ly.wgate); free_mat(ly.wup); free_mat(ly.wdown);
        cudaFree(ly.attn_norm); cudaFree(ly.ffn_norm); cudaFree(ly.bq); cudaFree(ly.bk); cudaFree(ly.bv); cudaFree(ly.bo);
    }
    if (m->tp > 1) m->logits = nullptr;   // lives inside the exchange window, freed with it below
    float *fs[] = {m->output_norm, m->gamma, m->x, m->xb, m->xb2, m->hb, m->hb2, m->q, m->k, m->v, m->logits, m->kc, m->vc, m->cos_t, m->sin_t};
    for (float *p : fs) if (p) cudaFree(p);
    int32_t *is[] = {m->gamma_map, m->d_token, m->d_pos, m->d_gen, m->d_gen_count, m->d_prompt, m->d_cursor};
    for (int32_t *p : is) if (p) cudaFree(p);
    if (m->tp > 1) {
        for (int r = 0; r < m->tp; r++) if (r != m->rank && m->tp_peers.win[r]) cudaIpcCloseMemHandle(m->tp_peers.win[r]);
        if (m->tp_win) cudaFree(m->tp_win);
        if (m->partial) cudaFree(m->partial);
        if (m->logits_local) cudaFree(m->logits_local);
        if (m->d_ar_epoch) cudaFree(m->d_ar_epoch);
        if (m->d_lg_epoch) cudaFree(m->d_lg_epoch);
    }
    if (m->d_trace) cudaFree(m->d_trace);
    if (m->d_trace2) cudaFree(m->d_trace2);
    if (m->pf_cap) { if (m->tp == 1) cudaFree(m->pf_x); cudaFree(m->pf_qkv); cudaFree(m->pf_g); cudaFree(m->pf_u); cudaFree(m->pf_hi); cudaFree(m->pf_lo); cudaFree(m->pf_split); }
    for (uint8_t *t : m->tile_bufs) if (t) cudaFree(t);
    for (void *q : {(void *)m->x_sh, (void *)m->qkv_sh, (void *)m->ao_sh, (void *)m->hb_sh, (void *)m->d_epoch, (void *)m->amax, (void *)m->arena}) if (q) cudaFree(q);
    if (m->qkv_bias) cudaFree(m->qkv_bias);
    if (m->d_tphases) cudaFree(m->d_tphases);
    if (m->d_bar) cudaFree(m->d_bar);
    if (m->part_acc) cudaFree(m->part_acc);
    if (m->part_ml) cudaFree(m->part_ml);
    if (m->h_stage) cudaFreeHost(m->h_stage);
    if (m->h_logits) cudaFreeHost(m->h_logits);
    if (m->sp_host) cudaFreeHost(m->sp_host);
    for (void *q : {(void *)m->sp_keys, (void *)m->sp_idx, (void *)m->sp_recent, (void *)m->sp_token}) if (q) cudaFree(q);
    if (m->ev0) cudaEventDestroy(m->ev0);
    if (m->ev1) cudaEventDestroy(m->ev1);
    if (m->st) cudaStreamDestroy(m->st);
    delete m;
}

static int ready(nl_model *m) {
    if (!m) return fail(NL_ERR_INVALID, "null model");
    if (!m->finalized) return fail(NL_ERR_STATE, "model not finalized");
    if (m->tp > 1 && !m->tp_ready) return fail(NL_ERR_STATE, "tensor-parallel model: nl_tp_import_handles has not been called");
    return set_dev(m);
}

int nl_forward_batch(nl_model *m, int32_t B, const int32_t *tokens, const int32_t *pos, float *logits_out) {
    int rc = ready(m); if (rc) return rc;
    if (B < 1 || B > m->B) return fail(NL_ERR_INVALID, "batch %d out of range [1,%d]", B, m->B);
    if (!tokens || !pos) return fail(NL_ERR_INVALID, "null argument");
    for (int b = 0; b < B; b++) {
        // the Go engine would panic on the slice index; report instead
        if (tokens[b] < 0 || tokens[b] >= m->c.vocab_size) return fail(NL_ERR_INVALID, "token %d out of range [0,%d)", tokens[b], m->c.vocab_size);
        if (pos[b] < 0 || pos[b] >= m->c.seq_len) return fail(NL_ERR_INVALID, "pos %d out of range [0,%d)", pos[b], m->c.seq_len);
    }
    rc = build_graphs(m, B); if (rc) return rc;
    memcpy(m->h_stage, tokens, B * 4); memcpy(m->h_stage + 64, pos, B * 4);
    NL_CUDA(cudaMemcpyAsync(m->d_token, m->h_stage, B * 4, cudaMemcpyHostToDevice, m->st));
    NL_CUDA(cudaMemcpyAsync(m->d_pos, m->h_stage + 64, B * 4, cudaMemcpyHostToDevice, m->st));
    NL_CUDA(cudaGraphLaunch(B == 1 ? next_graph(m, false) : m->g_fwd[B], m->st));
    if (logits_out) NL_CUDA(cudaMemcpyAsync(m->h_logits, m->logits, (size_t)B * m->c.vocab_size * 4, cudaMemcpyDeviceToHost, m->st));
    NL_CUDA(cudaStreamSynchronize(m->st));
    if (logits_out) memcpy(logits_out, m->h_logits, (size_t)B * m->c.vocab_size * 4);
    return NL_OK;
}

int nl_forward(nl_model *m, int32_t token, int32_t pos, float *logits_out) { return nl_forward_batch(m, 1, &token, &pos, logits_out); }

int nl_get_logits(nl_model *m, float *logits_out) {
    int rc = ready(m); if (rc) return rc;
    if (!logits_out) return fail(NL_ERR_INVALID, "null argument");
    NL_CUDA(cudaMemcpyAsync(m->h_logits, m->logits, (size_t)m->c.vocab_size * 4, cudaMemcpyDeviceToHost, m->st));
    NL_CUDA(cudaStreamSynchronize(m->st));
    memcpy(logits_out, m->h_logits, (size_t)m->c.vocab_size * 4);
    return NL_OK;
}

// One sampling step of Engine.Generate on the logits the last forward left on the device (go/main.go:177-197, :294-398).
int nl_sample(nl_model *m, float temperature, int32_t top_k, float top_p, float rep_penalty, const int32_t *recent, int32_t n_recent, float u,
              int32_t *token_out) {
    int rc = ready(m); if (rc) return rc;
    if (!token_out || n_recent < 0 || (n_recent > 0 && !recent)) return fail(NL_ERR_INVALID, "bad argument");
    if (n_recent > nl_model::SP_RECENT_CAP) return fail(NL_ERR_INVALID, "repetition window of %d tokens exceeds %d", n_recent, nl_model::SP_RECENT_CAP);
    if (temperature > 0.f && !(top_p < 1.0f) && top_k < 1) return fail(NL_ERR_INVALID, "top_k must be >= 1");
    if (temperature > 0.f && !(u >= 0.f && u < 1.f)) return fail(NL_ERR_INVALID, "u must be in [0, 1)");   // (ignored by the greedy step)
    const int vocab = m->c.vocab_size;
    if (!m->sp_keys) {
        NL_CUDA(cudaMalloc(&m->sp_keys, (size_t)vocab * 2 * sizeof(uint32_t)));
        NL_CUDA(cudaMalloc(&m->sp_idx, (size_t)vocab * 2 * sizeof(int32_t)));
        NL_CUDA(cudaMalloc(&m->sp_recent, (size_t)nl_model::SP_RECENT_CAP * sizeof(int32_t)));
        NL_CUDA(cudaMalloc(&m->sp_token, sizeof(int32_t)));
        NL_CUDA(cudaMallocHost(&m->sp_host, (size_t)(nl_model::SP_RECENT_CAP + 1) * sizeof(int32_t)));
    }
    if (n_recent > 0) {
        memcpy(m->sp_host + 1, recent, (size_t)n_recent * sizeof(int32_t));   // pinned staging: `recent` is caller memory
        NL_CUDA(cudaMemcpyAsync(m->sp_recent, m->sp_host + 1, (size_t)n_recent * sizeof(int32_t), cudaMemcpyHostToDevice, m->st));
    }
    SampleArgs a;
    a.logits = m->logits; a.vocab = vocab; a.recent = m->sp_recent; a.n_recent = n_recent; a.rep_penalty = rep_penalty;
    a.temp = temperature; a.top_k = top_k; a.top_p = top_p; a.u = u;
    a.keys0 = m->sp_keys; a.keys1 = m->sp_keys + vocab; a.idx0 = m->sp_idx; a.idx1 = m->sp_idx + vocab; a.token_out = m->sp_token;
    if (launch_sample(a, m->st)) return fail(NL_ERR_CUDA, "sample kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    NL_CUDA(cudaMemcpyAsync(m->sp_host, m->sp_token, sizeof(int32_t), cudaMemcpyDeviceToHost, m->st));
    NL_CUDA(cudaStreamSynchronize(m->st));
    *token_out = m->sp_host[0];
    return NL_OK;
}

int nl_reset(nl_model *m) {
    int rc = ready(m); if (rc) return rc;
    size_t n = (size_t)m->B * m->c.n_layers * m->c.seq_len * m->kvd * 4;
    NL_CUDA(cudaMemsetAsync(m->kc, 0, n, m->st));
    NL_CUDA(cudaMemsetAsync(m->vc, 0, n, m->st));
    NL_CUDA(cudaStreamSynchronize(m->st));
    return NL_OK;
}

// token-by-token prefill of sequence 0 on the device (feed kernel + forward graph per token)
static int prefill_sequential(nl_model *m, const int32_t *tokens, int n, int pos0) {
    if (n > m->prompt_cap) return fail(NL_ERR_INVALID, "prompt too long");
    NL_CUDA(cudaMemcpyAsync(m->d_prompt, tokens, (size_t)n * 4, cudaMemcpyHostToDevice, m->st));
    NL_CUDA(cudaMemsetAsync(m->d_cursor, 0, 4, m->st));
    NL_CUDA(cudaStreamSynchronize(m->st));  // tokens is caller memory: the copy must be done before we return or reuse it
    if (m->pf_time) NL_CUDA(cudaEventRecord(m->ev0, m->st));
    if (n > 1) { int rcw = set_lg_want(m, 0); if (rcw) return rcw; }   // only the last token's logits are ever read
    for (int i = 0; i < n; i++) {
        if (i == n - 1 && n > 1) { int rcw = set_lg_want(m, 1); if (rcw) return rcw; }
        feed_prompt_kernel<<<1, 1, 0, m->st>>>(m->d_prompt, m->d_cursor, pos0, m->d_token, m->d_pos);
        NL_CUDA(cudaGraphLaunch(next_graph(m, false), m->st));
    }
    m->pf_launches = n * (m->launches_fwd + 1);
    NL_CUDA(cudaGetLastError());
    return NL_OK;
}

// One-pass prefill on the tensor cores: every projection of the T prompt tokens is ONE tcgen05 GEMM (nl_gemm.cuh) instead of T GEMVs.
static bool prefill_gemm_ok(const nl_model *m, int n) {
    if (getenv("NL_NO_GEMM_PREFILL") || n < 16 || m->hd != 64) return false;
    if (m->tp > 1 && (getenv("NL_NO_TP_PREFILL") || n < m->tp)) return false;
    for (const Layer &ly : m->L)
        for (const DevMat *w : {&ly.wq, &ly.wk, &ly.wv, &ly.wo, &ly.wgate, &ly.wup, &ly.wdown})
            if (!gemm_eligible(*w)) return false;
    return true;
}
static int prefill_gemm(nl_model *m, const int32_t *tokens, int n, int pos0) {
    const nl_config &c = m->c;
    const int dim = m->dim, qdim = m->qdim, kvd = m->kvd, ffn = m->ffn, S = c.seq_len, ld = qdim + 2 * kvd;
    cudaStream_t st = m->st;
    const bool tpar = m->tp > 1;
    float *pf_part = tpar ? reinterpret_cast<float *>(m->tp_win + m->tp_lay.pf_part) : nullptr;
    { int rc0 = ensure_pf(m); if (rc0) return rc0; }
    NL_CUDA(cudaMemcpyAsync(m->d_prompt, tokens, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    NL_CUDA(cudaStreamSynchronize(st));   // tokens is caller memory
    if (m->pf_time) NL_CUDA(cudaEventRecord(m->ev0, st));
    {
        dim3 grid((dim + 255) / 256, n);
        embed_kernel<<<grid, 256, 0, st>>>(m->tok_embd, m->d_prompt, m->gamma, m->gamma_map, m->pf_x, dim);
    }
    m->pf_launches = 1 + c.n_layers * 7 + 1;   // embedding; per layer 2 norms, o and down GEMMs, RoPE/KV write, attention, SwiGLU split (q|k|v and gate|up: counted where they are launched); LM-head GEMV
    int rc;
    for (int l = 0; l < c.n_layers; l++) {
        Layer &ly = m->L[l];
        rmsnorm_split_kernel<<<n, 256, 0, st>>>(m->pf_x, ly.attn_norm, m->pf_hi, m->pf_lo, dim, c.rms_norm_eps, gemm_atile());
        {
            const GemmOut o[3] = {{&ly.wq, ly.bq, m->pf_qkv, ld, GEPI_STORE}, {&ly.wk, ly.bk, m->pf_qkv + qdim, ld, GEPI_STORE}, {&ly.wv, ly.bv, m->pf_qkv + qdim + kvd, ld, GEPI_STORE}};
            if ((rc = gemm_run_multi(o, 3, m->pf_hi, m->pf_lo, n, st, &m->pf_launches, m->pf_split))) return rc;
        }
        PrefillAttn a;
        a.qkv = m->pf_qkv; a.ld = ld; a.T = n; a.pos0 = pos0;
        a.kcache = m->kc + (int64_t)l * S * kvd; a.vcache = m->vc + (int64_t)l * S * kvd;
        a.cos_t = m->cos_t; a.sin_t = m->sin_t; a.out_hi = m->pf_hi; a.out_lo = m->pf_lo; a.tiled = gemm_atile();
        a.n_heads = m->nH; a.n_kv_heads = m->nKV; a.qk_norm = c.qk_norm; a.conj = c.rope_conjugate; a.eps = c.rms_norm_eps;
        a.scale = (float)(1.0 / sqrt((double)m->hd));
        rope_kv_kernel<<<n, 256, 0, st>>>(a);
        if (getenv("NL_PREFILL_ATTN_CC")) attn_prefill_kernel<<<dim3(m->nH, (n + 31) / 32), 256, 0, st>>>(a);   // (A/B: the CUDA-core kernel)
        else attn_prefill_tc_kernel<<<dim3(m->nH, (n + PA_QT - 1) / PA_QT), 256, 0, st>>>(a);
        if (tpar) {   // row-split o-projection: this rank's partial, then the reduce-scatter / all-gather of the rows (nl_tp.cuh)
            if ((rc = gemm_run(ly.wo, m->pf_hi, m->pf_lo, n, ly.bo, pf_part, dim, GEPI_STORE, st, m->pf_split))) return rc;
            tp_reduce_rows_kernel<<<m->opts.num_sms, 256, 0, st>>>(n, dim, m->tp_peers, m->tp_lay, m->rank, m->tp, m->d_pf_epoch);
            tp_rows_done_kernel<<<1, 32, 0, st>>>(m->tp_peers, m->tp_lay, m->rank, m->tp, m->d_pf_epoch);
            m->pf_launches += 2;
        } else if ((rc = gemm_run(ly.wo, m->pf_hi, m->pf_lo, n, ly.bo, m->pf_x, dim, GEPI_RESID, st, m->pf_split))) return rc;
        rmsnorm_split_kernel<<<n, 256, 0, st>>>(m->pf_x, ly.ffn_norm, m->pf_hi, m->pf_lo, dim, c.rms_norm_eps, gemm_atile());
        {
            const GemmOut o[2] = {{&ly.wgate, nullptr, m->pf_g, ffn, GEPI_STORE}, {&ly.wup, nullptr, m->pf_u, ffn, GEPI_STORE}};
            if ((rc = gemm_run_multi(o, 2, m->pf_hi, m->pf_lo, n, st, &m->pf_launches, m->pf_split))) return rc;
        }
        {
            const int64_t ne = (int64_t)n * ffn;
            swiglu_split_kernel<<<(unsigned)((ne / 2 + 255) / 256), 256, 0, st>>>(m->pf_g, m->pf_u, m->pf_hi, m->pf_lo, ne, ffn, gemm_atile());
        }
        if (tpar) {
            if ((rc = gemm_run(ly.wdown, m->pf_hi, m->pf_lo, n, nullptr, pf_part, dim, GEPI_STORE, st, m->pf_split))) return rc;
            tp_reduce_rows_kernel<<<m->opts.num_sms, 256, 0, st>>>(n, dim, m->tp_peers, m->tp_lay, m->rank, m->tp, m->d_pf_epoch);
            tp_rows_done_kernel<<<1, 32, 0, st>>>(m->tp_peers, m->tp_lay, m->rank, m->tp, m->d_pf_epoch);
            m->pf_launches += 2;
        } else if ((rc = gemm_run(ly.wdown, m->pf_hi, m->pf_lo, n, nullptr, m->pf_x, dim, GEPI_RESID, st, m->pf_split))) return rc;
    }
    // the reference computes the LM head at every prompt position and uses only the last (go/main.go:160-166): last row only
    NL_CUDA(cudaMemcpyAsync(m->x, m->pf_x + (size_t)(n - 1) * dim, (size_t)dim * 4, cudaMemcpyDeviceToDevice, st));
    const DevMat &out = tpar ? m->lm_view : (m->output.present() ? m->output : m->tok_embd);
    MatRef r = {&out, nullptr, nullptr, tpar ? m->logits_local : m->logits, tpar ? m->lvocab : c.vocab_size};
    rc = gemv_dispatch(&r, 1, m->x, dim, 1, EPI_STORE, m->output_norm, c.rms_norm_eps, m->xb, m->opts, st, nullptr);
    if (rc) return rc;
    if (tpar) {   // vocab-split LM head: every rank drops its shard into every window (m->logits points into this rank's window)
        tp_allgather_logits_kernel<<<1, 1024, 0, st>>>(m->logits_local, m->lvocab, m->tp_peers, m->tp_lay, m->rank, m->tp, m->d_lg_epoch);
        m->pf_launches++;
    }
    NL_CUDA(cudaGetLastError());
    return NL_OK;
}

static int prefill_check(nl_model *m, const int32_t *tokens, int32_t n, int32_t pos0) {
    if (!tokens || n <= 0) return fail(NL_ERR_INVALID, "empty prompt");
    if (pos0 < 0 || pos0 + n > m->c.seq_len) return fail(NL_ERR_INVALID, "positions [%d,%d) exceed seq_len %d", pos0, pos0 + n, m->c.seq_len);
    for (int i = 0; i < n; i++) if (tokens[i] < 0 || tokens[i] >= m->c.vocab_size) return fail(NL_ERR_INVALID, "token %d out of range", tokens[i]);
    return NL_OK;
}

int nl_prefill(nl_model *m, const int32_t *tokens, int32_t n, int32_t pos0, float *logits_last) {
    int rc = ready(m); if (rc) return rc;
    rc = prefill_check(m, tokens, n, pos0); if (rc) return rc;
    rc = prefill_gemm_ok(m, n) ? prefill_gemm(m, tokens, n, pos0) : prefill_sequential(m, tokens, n, pos0); if (rc) return rc;
    if (logits_last) NL_CUDA(cudaMemcpyAsync(m->h_logits, m->logits, (size_t)m->c.vocab_size * 4, cudaMemcpyDeviceToHost, m->st));
    NL_CUDA(cudaStreamSynchronize(m->st));
    if (logits_last) memcpy(logits_last, m->h_logits, (size_t)m->c.vocab_size * 4);
    return NL_OK;
}

int nl_bench_prefill(nl_model *m, const int32_t *tokens, int32_t n, int32_t pos0, float *ms_out, int32_t *launches_out) {
    int rc = ready(m); if (rc) return rc;
    if (!ms_out) return fail(NL_ERR_INVALID, "null argument");
    rc = prefill_check(m, tokens, n, pos0); if (rc) return rc;
    const bool gemm = prefill_gemm_ok(m, n);
    if (gemm) { rc = ensure_pf(m); if (rc) return rc; }
    m->pf_time = true;   // the prefill records ev0 after its token copy has landed
    rc = gemm ? prefill_gemm(m, tokens, n, pos0) : prefill_sequential(m, tokens, n, pos0);
    m->pf_time = false;
    if (rc) return rc;
    NL_CUDA(cudaEventRecord(m->ev1, m->st));
    NL_CUDA(cudaStreamSynchronize(m->st));
    NL_CUDA(cudaEventElapsedTime(ms_out, m->ev0, m->ev1));
    if (launches_out) *launches_out = m->pf_launches;
    return NL_OK;
}

int nl_generate_greedy(nl_model *m, const int32_t *prompt, int32_t n_prompt, int32_t n_new, int32_t eos_id, int32_t *out_tokens, int32_t *n_out) {
    int rc = ready(m); if (rc) return rc;
    if (!prompt || n_prompt <= 0 || !out_tokens || !n_out || n_new < 0) return fail(NL_ERR_INVALID, "bad argument");
    for (int i = 0; i < n_prompt; i++) if (prompt[i] < 0 || prompt[i] >= m->c.vocab_size) return fail(NL_ERR_INVALID, "token %d out of range", prompt[i]);
    *n_out = 0;
    const int S = m->c.seq_len;
    rc = nl_reset(m); if (rc) return rc;                         // go/main.go:156
    int n_fed = n_prompt < S - 1 ? n_prompt : S - 1;             // prefill stops at seq_len-1, go/main.go:160-166
    if (n_fed < 1) n_fed = 1;
    rc = prefill_sequential(m, prompt, n_fed, 0); if (rc) return rc;
    int pos = n_fed;
    NL_CUDA(cudaMemsetAsync(m->d_gen_count, 0, 4, m->st));
    rc = set_lg_want(m, 0); if (rc) return rc;   // the loop below feeds argmax pairs back on the device: no logits cross NVLink
    // each step: sample (argmax), then Forward(next, pos), pos++, stop when pos >= seq_len (go/main.go:173-218).
    // EOS is only visible on the host, so run in chunks and trim.
    int produced = 0; bool done = false;
    std::vector<int32_t> chunk(64);
    while (produced < n_new && !done) {
        int steps = n_new - produced < 64 ? n_new - produced : 64;
        int run = 0;
        for (int i = 0; i < steps; i++) {
            if (pos >= S) {  // the reference samples one last token, then Forward would overflow: it breaks after Forward at pos==S-1
                break;
            }
            NL_CUDA(cudaGraphLaunch(next_graph(m, true), m->st));
            pos++; run++;
        }
        if (run == 0) break;
        NL_CUDA(cudaMemcpyAsync(chunk.data(), m->d_gen + produced, (size_t)run * 4, cudaMemcpyDeviceToHost, m->st));
        NL_CUDA(cudaStreamSynchronize(m->st));
        for (int i = 0; i < run; i++) {
            out_tokens[produced++] = chunk[i];
            if (eos_id >= 0 && chunk[i] == eos_id) { done = true; break; }
        }
        if (pos >= S) done = true;
    }
    *n_out = produced;
    return set_lg_want(m, 1);
}

int nl_bench_decode(nl_model *m, int32_t token, int32_t pos0, int32_t n_steps, float *ms_out) {
    int rc = ready(m); if (rc) return rc;
    if (!ms_out || n_steps <= 0 || pos0 < 0 || pos0 + n_steps > m->c.seq_len || token < 0 || token >= m->c.vocab_size) return fail(NL_ERR_INVALID, "bad argument");
    // the first step is a plain forward of (token, pos0); every later step takes the argmax of the step before and the next position
    m->h_stage[0] = token; m->h_stage[64] = pos0;
    NL_CUDA(cudaMemcpyAsync(m->d_token, m->h_stage, 4, cudaMemcpyHostToDevice, m->st));
    NL_CUDA(cudaMemcpyAsync(m->d_pos, m->h_stage + 64, 4, cudaMemcpyHostToDevice, m->st));
    NL_CUDA(cudaMemsetAsync(m->d_gen_count, 0, 4, m->st));
    rc = set_lg_want(m, 0); if (rc) return rc;
    NL_CUDA(cudaEventRecord(m->ev0, m->st));
    NL_CUDA(cudaGraphLaunch(next_graph(m, false), m->st));
    for (int i = 1; i < n_steps; i++) NL_CUDA(cudaGraphLaunch(next_graph(m, true), m->st));
    NL_CUDA(cudaEventRecord(m->ev1, m->st));
    rc = set_lg_want(m, 1); if (rc) return rc;
    NL_CUDA(cudaStreamSynchronize(m->st));
    NL_CUDA(cudaEventElapsedTime(ms_out, m->ev0, m->ev1));
    if (m->d_trace && getenv("NL_TRACE")) {
        const int tg = m->tile_grid, tp_ = m->targs.n_phases;
        size_t n = (size_t)tg * tp_ * 8;
        std::vector<unsigned long long> h(n);
        NL_CUDA(cudaMemcpy(h.data(), m->d_trace, n * 8, cudaMemcpyDeviceToHost));
        // tensor parallel: one file per rank (<NL_TRACE>.r<rank>)
        const std::string p1 = std::string(getenv("NL_TRACE")) + (m->tp > 1 ? ".r" + std::to_string(m->rank) : std::string());
        FILE *f = fopen(p1.c_str(), "wb");
        if (f) { int hdr[2] = {tg, tp_}; fwrite(hdr, 4, 2, f); fwrite(h.data(), 8, n, f); fclose(f); }
        if (m->d_trace2 && m->tile_ok) {   // clock64 sub-stamps -> <NL_TRACE>.ck
            std::vector<unsigned long long> h2(n * 2);
            NL_CUDA(cudaMemcpy(h2.data(), m->d_trace2, n * 16, cudaMemcpyDeviceToHost));
            std::string p2 = p1 + ".ck";
            FILE *f2 = fopen(p2.c_str(), "wb");
            if (f2) { int hdr2[2] = {tg, tp_}; fwrite(hdr2, 4, 2, f2); fwrite(h2.data(), 8, n * 2, f2); fclose(f2); }
        }
    }
    return NL_OK;
}

int nl_launches_per_token(const nl_model *m) { return m ? m->launches_fwd + 1 : 0; }
int64_t nl_weight_bytes(const nl_model *m) { return m ? m->weight_bytes : 0; }
const char *nl_decode_path(const nl_model *m) {
    if (!m) return "";
    return m->tile_ok ? "decode_tiled_kernel" : "gemv_stream_kernel chain";
}

// ---- operator-level hooks ----
int nl_dequant(uint32_t type, const void *host_src, int64_t n, float *host_dst) {
    if (!host_src || !host_dst || n <= 0) return fail(NL_ERR_INVALID, "bad argument");
    if (!type_supported((int)type)) return fail(NL_ERR_UNSUPPORTED, "unsupported tensor type %u", type);
    int be = blk_elems((int)type);
    if (n % be) return fail(NL_ERR_INVALID, "n=%lld not a multiple of the block size %d", (long long)n, be);
    int dev = 0; cudaGetDevice(&dev);
    int rc = check_device(dev); if (rc) return rc;
    // present the data as a [n/be', be'] matrix so the F16/F32 unit constraints hold for any n
    DevMat m;
    int64_t cols = be == 1 ? 1 : be; int64_t rows = n / cols;
    m.type = (int)type;
    size_t nbytes = (size_t)tensor_nbytes((int)type, n);
    if (type == NL_F16 || type == NL_F32) {  // bypass the 16-byte-unit check of upload_mat (no GEMV on this path)
        m.rows = rows; m.cols = cols; m.qs_bytes = nbytes;
        NL_CUDA(cudaMalloc(&m.qs, nbytes));
        NL_CUDA(cudaMemcpy(m.qs, host_src, nbytes, cudaMemcpyHostToDevice));
    } else {
        rc = upload_mat(m, (int)type, rows, cols, host_src, nbytes, 0); if (rc) return rc;
    }
    float *out = nullptr;
    cudaError_t e = cudaMalloc(&out, (size_t)n * 4);
    if (e != cudaSuccess) { free_mat(m); return fail(NL_ERR_OOM, "cudaMalloc: %s", cudaGetErrorString(e)); }
    rc = dequant_mat(m, out, 0);
    if (!rc) { e = cudaMemcpy(host_dst, out, (size_t)n * 4, cudaMemcpyDeviceToHost); if (e != cudaSuccess) rc = fail(NL_ERR_CUDA, "D2H: %s", cudaGetErrorString(e)); }
    cudaFree(out); free_mat(m);
    return rc;
}

struct nl_matrix {
    int device = 0;
    std::vector<DevMat> copies;  // replicas for L2-cold benchmarking; [0] is the matrix
    float *x = nullptr, *out = nullptr; int xcap = 0;
    __nv_bfloat16 *xh = nullptr, *xl = nullptr;   // bf16 planes of x for the tensor-core path
    float *split = nullptr;                       // split-K partial sums (nl_gemm2.cuh)
    cudaStream_t st = nullptr;
    GemvOpts opts{148, true, true};
    // batch-1 Q4_0: fragment-tiled replicas + one-phase descriptors for the tensor-core GEMV (nl_tile.cuh)
    std::vector<uint8_t *> tiles;
    TilePhase *d_tph = nullptr; unsigned int *d_tbar = nullptr; int tph_cap = 0;
};

// tiled replica of copy `idx` (built on first use) and a one-phase descriptor per replica
static int matrix_tiles(nl_matrix *w, int n_copies) {
    const DevMat &m0 = w->copies[0];
    if (getenv("NL_NO_TILED") || !tile_eligible(m0)) return 1;
    while ((int)w->tiles.size() < n_copies) {
        uint8_t *t = nullptr;
        const DevMat *one[1] = {&w->copies[w->tiles.size()]};
        int rc = make_tiles(&t, one, 1, false, w->st); if (rc) return rc;
        w->tiles.push_back(t);
    }
    if (w->tph_cap < n_copies) {
        if (w->d_tph) cudaFree(w->d_tph);
        if (!w->d_tbar) { NL_CUDA(cudaMalloc(&w->d_tbar, 4)); NL_CUDA(cudaMemset(w->d_tbar, 0, 4)); }
        std::vector<TilePhase> ph(n_copies);
        for (int i = 0; i < n_copies; i++)
            tile_gemv_phase(ph[i], w->tiles[i], (int)(m0.rows / 16), (int)m0.cols, 1, (int)m0.rows, TEPI_STORE, w->x, 0, nullptr, nullptr, w->out, 0);
        NL_CUDA(cudaMalloc(&w->d_tph, n_copies * sizeof(TilePhase)));
        NL_CUDA(cudaMemcpy(w->d_tph, ph.data(), n_copies * sizeof(TilePhase), cudaMemcpyHostToDevice));
        w->tph_cap = n_copies;
    }
    return NL_OK;
}
static int matrix_tiled_gemv(nl_matrix *w, int idx) {
    TileArgs a; memset(&a, 0, sizeof a);
    a.phases = w->d_tph + idx; a.n_phases = 1; a.bar = w->d_tbar; a.inflight = tile_inflight(); a.att_chunk = 96;
    const int units = (int)(w->copies[0].rows / 16);
    if (launch_tiled(w->copies[0].type, a, units < w->opts.num_sms ? units : w->opts.num_sms, w->st)) return fail(NL_ERR_CUDA, "tiled gemv launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    return NL_OK;
}

int nl_matrix_create(uint32_t type, const void *host_w, int64_t rows, int64_t cols, int32_t device, nl_matrix **out) {
    if (!host_w || !out) return fail(NL_ERR_INVALID, "null argument");
    *out = nullptr;
    int rc = check_device(device); if (rc) return rc;
    if (rows > INT32_MAX || cols > INT32_MAX) return fail(NL_ERR_INVALID, "matrix too large");
    nl_matrix *w = new nl_matrix(); w->device = device; w->copies.resize(1); w->opts = default_opts(device);
    cudaStreamCreateWithFlags(&w->st, cudaStreamNonBlocking);
    int64_t nb = tensor_nbytes((int)type, rows * cols);
    rc = upload_mat(w->copies[0], (int)type, rows, cols, host_w, nb < 0 ? 0 : (size_t)nb, w->st);
    if (rc) { nl_matrix_destroy(w); return rc; }
    *out = w;
    return NL_OK;
}

static int matrix_buffers(nl_matrix *w, int batch) {
    if (batch <= w->xcap) return NL_OK;
    if (w->x) cudaFree(w->x); if (w->out) cudaFree(w->out);
    if (w->xh) cudaFree(w->xh); if (w->xl) cudaFree(w->xl);
    w->x = w->out = nullptr; w->xh = w->xl = nullptr;
    w->tph_cap = 0;   // the one-phase descriptors point at x / out
    const size_t bt = ((size_t)batch + 127) & ~(size_t)127;   // (whole 128-row tiles, see ensure_pf)
    NL_CUDA(cudaMalloc(&w->xh, bt * w->copies[0].cols * 2));
    NL_CUDA(cudaMalloc(&w->xl, bt * w->copies[0].cols * 2));
    if (!w->split) NL_CUDA(cudaMalloc(&w->split, G2_SPLIT_BYTES));
    NL_CUDA(cudaMalloc(&w->x, (size_t)batch * w->copies[0].cols * 4));
    NL_CUDA(cudaMalloc(&w->out, (size_t)batch * w->copies[0].rows * 4));
    w->xcap = batch;
    return NL_OK;
}

int nl_matrix_matmul(nl_matrix *w, const float *host_x, int32_t batch, float *host_out) {
    if (!w || !host_x || !host_out || batch < 1 || batch > 4096) return fail(NL_ERR_INVALID, "bad argument");
    NL_CUDA(cudaSetDevice(w->device));
    int rc = matrix_buffers(w, batch); if (rc) return rc;
    const DevMat &m = w->copies[0];
    NL_CUDA(cudaMemcpyAsync(w->x, host_x, (size_t)batch * m.cols * 4, cudaMemcpyHostToDevice, w->st));
    const int gemm_min = getenv("NL_GEMM_MIN_BATCH") ? atoi(getenv("NL_GEMM_MIN_BATCH")) : 16;
    if (batch >= gemm_min && gemm_eligible(m)) {   // many rows of x: the T-token GEMM on the tensor cores
        rc = split_planes(w->x, w->xh, w->xl, (int64_t)batch * m.cols, (int)m.cols, w->st); if (rc) return rc;
        rc = gemm_run(m, w->xh, w->xl, batch, nullptr, w->out, (int)m.rows, GEPI_STORE, w->st, w->split); if (rc) return rc;
    } else if (batch == 1 && (rc = matrix_tiles(w, 1)) <= 0) {   // batch 1, Q4_0: the tensor-core GEMV of the decode path
        if (rc) return rc;
        rc = matrix_tiled_gemv(w, 0); if (rc) return rc;
    } else {
        if (batch > 64) return fail(NL_ERR_INVALID, "batch %d > 64 needs a Q4_0/Q8_0/F16 matrix with cols %% 32 == 0", batch);
        MatRef r = {&m, nullptr, nullptr, w->out, (int)m.rows};
        rc = gemv_dispatch(&r, 1, w->x, (int)m.cols, batch, EPI_STORE, nullptr, 0.f, nullptr, w->opts, w->st, nullptr); if (rc) return rc;
    }
    NL_CUDA(cudaMemcpyAsync(host_out, w->out, (size_t)batch * m.rows * 4, cudaMemcpyDeviceToHost, w->st));
    NL_CUDA(cudaStreamSynchronize(w->st));
    return NL_OK;
}

int nl_matrix_bench(nl_matrix *w, int32_t batch, int32_t n_copies, int32_t warmup, int32_t iters, float *ms_out) {
    if (!w || !ms_out || batch < 1 || batch > 4096 || n_copies < 1 || iters < 1) return fail(NL_ERR_INVALID, "bad argument");
    NL_CUDA(cudaSetDevice(w->device));
    int rc = matrix_buffers(w, batch); if (rc) return rc;
    const DevMat src = w->copies[0];
    const int gemm_min = getenv("NL_GEMM_MIN_BATCH") ? atoi(getenv("NL_GEMM_MIN_BATCH")) : 16;
    if (batch > 64 && !(batch >= gemm_min && gemm_eligible(src))) return fail(NL_ERR_INVALID, "batch %d > 64 needs a Q4_0/Q8_0/F16 matrix with cols %% 32 == 0", batch);
    while ((int)w->copies.size() < n_copies) {
        DevMat c = src; c.qs = nullptr; c.d = nullptr;
        NL_CUDA(cudaMalloc(&c.qs, src.qs_bytes));
        NL_CUDA(cudaMemcpy(c.qs, src.qs, src.qs_bytes, cudaMemcpyDeviceToDevice));
        if (src.d) { NL_CUDA(cudaMalloc(&c.d, src.d_bytes)); NL_CUDA(cudaMemcpy(c.d, src.d, src.d_bytes, cudaMemcpyDeviceToDevice)); }
        w->copies.push_back(c);
    }
    NL_CUDA(cudaMemsetAsync(w->x, 0, (size_t)batch * src.cols * 4, w->st));
    cudaEvent_t e0, e1; NL_CUDA(cudaEventCreate(&e0)); NL_CUDA(cudaEventCreate(&e1));
    const int tiled = batch == 1 ? matrix_tiles(w, n_copies) : 1;
    if (tiled < 0) return tiled;
    const bool gemm = batch >= gemm_min && gemm_eligible(src);   // the T-row GEMM on the tensor cores (the activation planes are split once, outside the timed loop)
    if (gemm) { rc = split_planes(w->x, w->xh, w->xl, (int64_t)batch * src.cols, (int)src.cols, w->st); if (rc) return rc; }
    int idx = 0;
    for (int i = 0; i < warmup + iters; i++) {
        if (i == warmup) NL_CUDA(cudaEventRecord(e0, w->st));
        if (gemm) { rc = gemm_run(w->copies[idx], w->xh, w->xl, batch, nullptr, w->out, (int)src.rows, GEPI_STORE, w->st, w->split); if (rc) return rc; }
        else if (tiled == 0) { rc = matrix_tiled_gemv(w, idx); if (rc) return rc; }
        else {
            MatRef r = {&w->copies[idx], nullptr, nullptr, w->out, (int)src.rows};
            rc = gemv_dispatch(&r, 1, w->x, (int)src.cols, batch, EPI_STORE, nullptr, 0.f, nullptr, w->opts, w->st, nullptr); if (rc) return rc;
        }
        idx = (idx + 1) % n_copies;
    }
    NL_CUDA(cudaEventRecord(e1, w->st));
    NL_CUDA(cudaStreamSynchronize(w->st));
    float ms = 0; NL_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *ms_out = ms / iters;
    return NL_OK;
}

void nl_matrix_destroy(nl_matrix *w) {
    if (!w) return;
    cudaSetDevice(w->device);
    if (w->st) cudaStreamSynchronize(w->st);
    for (auto &c : w->copies) free_mat(c);
    for (uint8_t *t : w->tiles) if (t) cudaFree(t);
    if (w->d_tph) cudaFree(w->d_tph);
    if (w->d_tbar) cudaFree(w->d_tbar);
    if (w->x) cudaFree(w->x); if (w->out) cudaFree(w->out);
    if (w->xh) cudaFree(w->xh); if (w->xl) cudaFree(w->xl);
    if (w->split) cudaFree(w->split);
    if (w->st) cudaStreamDestroy(w->st);
    delete w;
}

int nl_matmul(uint32_t type, const void *host_w, int64_t rows, int64_t cols, const float *host_x, int32_t batch, float *host_out) {
    int dev = 0; cudaGetDevice(&dev);
    nl_matrix *w = nullptr;
    int rc = nl_matrix_create(type, host_w, rows, cols, dev, &w); if (rc) return rc;
    rc = nl_matrix_matmul(w, host_x, batch, host_out);
    nl_matrix_destroy(w);
    return rc;
}

int nl_tp_export_handle(nl_model *m, void *handle64) {
    if (!m || !handle64) return fail(NL_ERR_INVALID, "null argument");
    if (m->tp <= 1) return fail(NL_ERR_STATE, "model is not tensor-parallel (tp_size == 1)");
    if (!m->finalized || !m->tp_win) return fail(NL_ERR_STATE, "nl_finalize first");
    int rc = set_dev(m); if (rc) return rc;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle is 64 bytes");
    cudaIpcMemHandle_t hnd;
    NL_CUDA(cudaIpcGetMemHandle(&hnd, m->tp_win));
    memcpy(handle64, &hnd, 64);
    return NL_OK;
}

int nl_tp_import_handles(nl_model *m, const void *handles64_by_rank, int32_t n_ranks) {
    if (!m || !handles64_by_rank) return fail(NL_ERR_INVALID, "null argument");
    if (m->tp <= 1) return fail(NL_ERR_STATE, "model is not tensor-parallel (tp_size == 1)");
    if (n_ranks != m->tp) return fail(NL_ERR_INVALID, "%d handles for tp_size %d", n_ranks, m->tp);
    if (!m->finalized || !m->tp_win) return fail(NL_ERR_STATE, "nl_finalize first");
    if (m->tp_ready) return fail(NL_ERR_STATE, "peer windows already imported");
    int rc = set_dev(m); if (rc) return rc;
    for (int r = 0; r < m->tp; r++) {
        if (r == m->rank) { m->tp_peers.win[r] = m->tp_win; continue; }
        cudaIpcMemHandle_t hnd;
        memcpy(&hnd, (const uint8_t *)handles64_by_rank + 64 * r, 64);
        void *p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, hnd, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) return fail(NL_ERR_CUDA, "cudaIpcOpenMemHandle(rank %d): %s", r, cudaGetErrorString(e));
        m->tp_peers.win[r] = (uint8_t *)p;
    }
    m->tp_ready = true;
    rc = build_tiled(m);   // the shards as tiles, exchange inside the persistent kernel (falls back to the per-matrix chain when not eligible)
    if (rc) { m->tp_ready = false; return rc; }
    rc = build_graphs(m, 1);
    if (rc) { m->tp_ready = false; return rc; }
    return NL_OK;
}

}  // extern "C"
