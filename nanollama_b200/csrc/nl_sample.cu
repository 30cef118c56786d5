// nl_sample.cu — one sampling step of Engine.Generate on the device (go/main.go:177-197 repetition penalty, :294-343 sampleTopK,
// :346-398 sampleTopP, :400-408 argmax), so that the 384 KB logits vector of a 96k vocabulary never leaves HBM and the host does
// not sort it (the reference sorts the whole vocabulary per token in sampleTopP).
//
// One CTA: (1) repetition penalty in place; (2) a STABLE descending order of the vocabulary by logit -- LSD radix sort, 8 passes of
// 4 bits over order-preserving keys, every thread owning a contiguous chunk so that equal logits keep their index order (the
// reference's insertion order in sampleTopK; a stable sort in sampleTopP); (3) the reference's arithmetic on that order, with its
// SEQUENTIAL fp32 sums (a running cdf decides which token a random number selects, so the order of the additions is part of the
// result): weights of 1024 sorted positions are evaluated in parallel into shared memory, one thread adds them up in order.
// Deviations from the reference: the normalising sum of sampleTopP over the whole vocabulary is a fixed-order parallel sum (the
// reference adds 96k terms in index order), so probabilities can differ in the last bit; the random number is supplied by the host
// (rng.Float32()), which keeps the stream of random numbers the host's.
#include "nl_sample.cuh"

namespace nl {

constexpr int SP_T = 1024;          // threads
constexpr int SP_BINS = 16;         // 4-bit digits
constexpr int SP_PASSES = 8;

// ascending order of the key = descending order of the value
__device__ __forceinline__ uint32_t desc_key(float f) {
    const uint32_t b = __float_as_uint(f);
    const uint32_t asc = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
    return ~asc;
}

__device__ __forceinline__ int block_excl_scan(int v, int *warp_tot, int tid) {   // exclusive prefix of v over the block
    const int lane = tid & 31, warp = tid >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int t = warp_tot[lane], ti = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, ti, o);
            if (lane >= o) ti += n;
        }
        warp_tot[lane] = ti - t;   // exclusive prefix of the warp totals
    }
    __syncthreads();
    const int r = warp_tot[warp] + inc - v;
    __syncthreads();   // warp_tot is reused by the next call
    return r;
}

// float32(math.Exp(float64((x - mx) / temp))), go/main.go:332 / :370
__device__ __forceinline__ float soft_weight(float x, float mx, float temp) { return (float)exp((double)__fdiv_rn(__fsub_rn(x, mx), temp)); }

__global__ void __launch_bounds__(SP_T, 1) sample_kernel(const SampleArgs A) {
    extern __shared__ int cnt[];   // [SP_BINS][SP_T] digit counts of the radix passes
    __shared__ float wbuf[SP_T];
    __shared__ int warp_tot[32];
    __shared__ float red_v[32];
    __shared__ int red_i[32];
    __shared__ int s_found, s_pick;
    __shared__ float s_cum, s_bcast;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = A.vocab;
    float *lg = A.logits;

    // ---- repetition penalty (go/main.go:177-187): once per occurrence in `recent`; the thread of a token's FIRST occurrence applies all of them
    if (A.rep_penalty > 1.0f && A.n_recent > 0) {
        for (int i = tid; i < A.n_recent; i += SP_T) {
            const int tok = A.recent[i];
            if (tok < 0 || tok >= n) continue;
            bool first = true;
            int c = 0;
            for (int j = 0; j < A.n_recent; j++)
                if (A.recent[j] == tok) { if (j < i) first = false; c++; }
            if (first) {
                float x = lg[tok];
                for (int k = 0; k < c; k++) x = x > 0.f ? __fdiv_rn(x, A.rep_penalty) : __fmul_rn(x, A.rep_penalty);
                lg[tok] = x;
            }
        }
        __syncthreads();
    }

    // ---- temp <= 0: argmax, first maximum (go/main.go:299-301, :400-408)
    if (!(A.temp > 0.f)) {
        float bv = -INFINITY;
        int bi = 0x7fffffff;
        for (int i = tid; i < n; i += SP_T) {
            const float v = lg[i];
            if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) { red_v[warp] = bv; red_i[warp] = bi; }
        __syncthreads();
        if (warp == 0) {
            bv = red_v[lane]; bi = red_i[lane];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            }
            if (lane == 0) *A.token_out = bi == 0x7fffffff ? 0 : bi;
        }
        return;
    }

    // ---- stable descending order by logit: keys + indices, LSD radix sort
    for (int i = tid; i < n; i += SP_T) { A.keys0[i] = desc_key(lg[i]); A.idx0[i] = i; }
    __syncthreads();
    const int chunk = (n + SP_T - 1) / SP_T;
    const int lo = min(n, tid * chunk), hi = min(n, lo + chunk);
    uint32_t *ks = A.keys0, *kd = A.keys1;
    int32_t *is = A.idx0, *id = A.idx1;
    for (int pass = 0; pass < SP_PASSES; pass++) {
        const int shift = 4 * pass;
        for (int d = 0; d < SP_BINS; d++) cnt[d * SP_T + tid] = 0;
        for (int i = lo; i < hi; i++) cnt[((ks[i] >> shift) & 15u) * SP_T + tid]++;
        __syncthreads();
        // exclusive scan of the flattened [digit][thread] array: 16 consecutive entries per thread, then across the block
        int s = 0;
        for (int j = 0; j < SP_BINS; j++) { const int v = cnt[SP_BINS * tid + j]; cnt[SP_BINS * tid + j] = s; s += v; }
        const int base = block_excl_scan(s, warp_tot, tid);
        for (int j = 0; j < SP_BINS; j++) cnt[SP_BINS * tid + j] += base;
        __syncthreads();
        for (int i = lo; i < hi; i++) {
            const uint32_t k = ks[i];
            const int pos = cnt[((k >> shift) & 15u) * SP_T + tid]++;
            kd[pos] = k; id[pos] = is[i];
        }
        __syncthreads();
        uint32_t *tk = ks; ks = kd; kd = tk;
        int32_t *ti = is; is = id; id = ti;
    }
    // (an even number of passes: the order is back in keys0 / idx0)
    const int32_t *ord = is;

    const bool topp = A.top_p < 1.0f;
    const float mx = lg[ord[0]];
    float inv_sum = 1.f;
    int limit;
    if (topp) {
        // probabilities over the whole vocabulary (go/main.go:368-378): fixed-order parallel sum
        float part = 0.f;
        for (int i = tid; i < n; i += SP_T) part = __fadd_rn(part, soft_weight(lg[i], mx, A.temp));
        part = warp_sum(part);
        if (lane == 0) red_v[warp] = part;
        __syncthreads();
        if (warp == 0) {
            float t = warp_sum(red_v[lane]);
            if (lane == 0) s_bcast = t;
        }
        __syncthreads();
        inv_sum = __fdiv_rn(1.0f, s_bcast);
        limit = n;
    } else {
        limit = A.top_k < n ? A.top_k : n;
    }

    // ---- pass A: running sum over the sorted order; top-p stops at the first position where it reaches top_p (:384-387), top-k
    // adds up its k weights (:328-334)
    if (tid == 0) { s_found = -1; s_pick = -1; s_cum = 0.f; }
    __syncthreads();
    for (int base = 0; base < limit; base += SP_T) {
        const int i = base + tid;
        float w = 0.f;
        if (i < limit) { w = soft_weight(lg[ord[i]], mx, A.temp); if (topp) w = __fmul_rn(w, inv_sum); }
        wbuf[tid] = w;
        __syncthreads();
        if (tid == 0) {
            float cum = s_cum;
            const int m = min(SP_T, limit - base);
            for (int j = 0; j < m; j++) {
                cum = __fadd_rn(cum, wbuf[j]);
                if (topp && cum >= A.top_p) { s_found = base + j; break; }
            }
            s_cum = cum;
        }
        __syncthreads();
        if (s_found >= 0) break;
    }
    int found = topp ? s_found : limit - 1;
    if (found < 0) {   // the probabilities never reach top_p (rounding): the reference returns the most likely token (:397)
        if (tid == 0) *A.token_out = ord[0];
        return;
    }
    const float r = __fmul_rn(A.u, s_cum);   // rng.Float32() * cumsum (:339, :388)
    __syncthreads();
    if (tid == 0) s_cum = 0.f;
    __syncthreads();
    // ---- pass B: the first position whose running cdf reaches r (:340-345, :389-395)
    for (int base = 0; base <= found; base += SP_T) {
        const int i = base + tid;
        float w = 0.f;
        if (i <= found) { w = soft_weight(lg[ord[i]], mx, A.temp); if (topp) w = __fmul_rn(w, inv_sum); }
        wbuf[tid] = w;
        __syncthreads();
        if (tid == 0) {
            float cdf = s_cum;
            const int m = min(SP_T, found + 1 - base);
            for (int j = 0; j < m; j++) {
                cdf = __fadd_rn(cdf, wbuf[j]);
                if (r <= cdf) { s_pick = base + j; break; }
            }
            s_cum = cdf;
        }
        __syncthreads();
        if (s_pick >= 0) break;
    }
    if (tid == 0) *A.token_out = ord[s_pick >= 0 ? s_pick : 0];
}

int launch_sample(const SampleArgs &a, cudaStream_t st) {
    static bool configured = false;
    const int smem = SP_BINS * SP_T * (int)sizeof(int);
    if (!configured) {
        if (cudaFuncSetAttribute(sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return -2;
        configured = true;
    }
    sample_kernel<<<1, SP_T, smem, st>>>(a);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

}  // namespace nl
