// nl_sample.cu — one sampling step of Engine.Generate on the device (go/main.go:177-197 repetition penalty, :294-343 sampleTopK,
// :346-398 sampleTopP, :400-408 argmax), so that the 384 KB logits vector of a 96k vocabulary never leaves HBM and the host does
// not sort it (the reference sorts the whole vocabulary per token in sampleTopP).
//
// One CTA: (1) repetition penalty in place; (2) a STABLE descending order by logit of the tokens that can matter -- the candidates
// that share the top 12 key bits of the `top_k` (top-p: 2048) largest logits, picked by a histogram and compacted in index order,
// then an LSD radix sort (8 passes of 4 bits over order-preserving keys, every thread owning a contiguous chunk, so equal logits keep
// their index order: the reference's insertion order in sampleTopK, a stable sort in sampleTopP); the whole vocabulary is ordered
// instead when the candidates are too many or top-p's nucleus outgrows them; (3) the reference's arithmetic on that order, with its
// SEQUENTIAL fp32 sums (a running cdf decides which token a random number selects, so the order of the additions is part of the
// result): 1024 weights at a time are evaluated in parallel into shared memory, one thread adds them up in order -- including
// sampleTopP's normalising sum over the whole vocabulary in index order.  The random number is supplied by the host
// (rng.Float32()), which keeps the stream of random numbers the host's.
// What can differ from the reference is the order among EQUAL entries only: sampleTopP sorts by normalised probability with an
// unstable sort (sort.Slice), so tokens whose probabilities round to the same fp32 value come in an arbitrary order there; here they
// come by logit, then by index (tests/test_sampler_algorithm.py pins this on the CPU).
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

// Stable LSD radix order of n (key, index) pairs, ascending keys; 8 passes of 4 bits, thread t owns the contiguous chunk t of the
// source.  Ping-pongs between (k0, i0) and (k1, i1); the result is back in (k0, i0).
__device__ void radix_order(uint32_t *k0, int32_t *i0, uint32_t *k1, int32_t *i1, int n, int *cnt, int *warp_tot, int tid) {
    const int chunk = (n + SP_T - 1) / SP_T;
    const int lo = min(n, tid * chunk), hi = min(n, lo + chunk);
    uint32_t *ks = k0, *kd = k1;
    int32_t *is = i0, *id = i1;
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
}

constexpr int SP_HBITS = 12, SP_HBINS = 1 << SP_HBITS;   // candidate selection: histogram of the top 12 key bits
constexpr int SP_CAND_MAX = 16384;                       // more candidates than this: order the whole vocabulary instead
constexpr int SP_CAND_TOPP = 2048;                       // top-p: first guess at the size of the nucleus

__global__ void __launch_bounds__(SP_T, 1) sample_kernel(const SampleArgs A) {
    extern __shared__ int cnt[];   // [SP_BINS][SP_T] digit counts of the radix passes; the candidate histogram [SP_HBINS] before them
    __shared__ __align__(16) float wbuf[SP_T];
    __shared__ int warp_tot[32];
    __shared__ float red_v[32];
    __shared__ int red_i[32];
    __shared__ int s_found, s_pick, s_bstar, s_ncand;
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

    // ---- the maximum and its first index (go/main.go:400-408; maxVal of :353-358 and top[0] of :326)
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
        if (lane == 0) { s_bcast = bv; s_pick = bi == 0x7fffffff ? 0 : bi; }
    }
    __syncthreads();
    const float mx = s_bcast;
    const int arg_max = s_pick;
    __syncthreads();
    if (!(A.temp > 0.f)) {   // temp <= 0: argmax (go/main.go:299-301)
        if (tid == 0) *A.token_out = arg_max;
        return;
    }

    const bool topp = A.top_p < 1.0f;
    float inv_sum = 1.f;
    if (topp) {
        // probabilities over the whole vocabulary (go/main.go:368-378).  The normalising sum is the reference's: fp32, in index order
        // (1024 weights at a time evaluated in parallel, one thread adds them up) -- a cdf over tens of thousands of near-equal
        // probabilities selects a different token when the sum differs in its last bit
        if (tid == 0) s_cum = 0.f;
        for (int base = 0; base < n; base += SP_T) {
            const int i = base + tid;
            wbuf[tid] = i < n ? soft_weight(lg[i], mx, A.temp) : 0.f;
            __syncthreads();
            if (tid == 0) {
                float acc = s_cum;
                const int m = min(SP_T, n - base);
                int j = 0;
                for (; j + 4 <= m; j += 4) {
                    const float4 w4 = *reinterpret_cast<const float4 *>(&wbuf[j]);
                    acc = __fadd_rn(acc, w4.x); acc = __fadd_rn(acc, w4.y); acc = __fadd_rn(acc, w4.z); acc = __fadd_rn(acc, w4.w);
                }
                for (; j < m; j++) acc = __fadd_rn(acc, wbuf[j]);
                s_cum = acc;
            }
            __syncthreads();
        }
        inv_sum = __fdiv_rn(1.0f, s_cum);
    }
    const int want = topp ? min(n, SP_CAND_TOPP) : min(n, A.top_k);

    // Two attempts at most: (0) order only the candidates that share the top key bits of the `want` largest logits (everything else is
    // smaller than all of them); (1) order the whole vocabulary -- when the candidates are too many (flat logits) or the nucleus of
    // top-p turns out to be larger than the candidate set.
    for (int attempt = (want > SP_CAND_MAX ? 1 : 0); attempt < 2; attempt++) {
        int n_ord;
        if (attempt == 0) {
            int *hist = cnt;
            for (int b = tid; b < SP_HBINS; b += SP_T) hist[b] = 0;
            if (tid == 0) { s_bstar = SP_HBINS - 1; s_ncand = n; }
            __syncthreads();
            for (int i = tid; i < n; i += SP_T) atomicAdd(&hist[desc_key(lg[i]) >> (32 - SP_HBITS)], 1);
            __syncthreads();
            // first bin whose inclusive prefix count reaches `want`
            constexpr int PER = SP_HBINS / SP_T;
            int h[PER], s = 0;
            for (int j = 0; j < PER; j++) { h[j] = hist[PER * tid + j]; s += h[j]; }
            int pre = block_excl_scan(s, warp_tot, tid);
            for (int j = 0; j < PER; j++) {
                if (pre < want && pre + h[j] >= want) { s_bstar = PER * tid + j; s_ncand = pre + h[j]; }
                pre += h[j];
            }
            __syncthreads();
            const int bstar = s_bstar;
            n_ord = s_ncand;
            __syncthreads();
            if (n_ord > SP_CAND_MAX) continue;   // (uniform) too many share the boundary bin: order everything
            // stable compaction of the candidates into (keys1, idx1): contiguous chunk per thread, index order kept
            const int chunk = (n + SP_T - 1) / SP_T;
            const int lo = min(n, tid * chunk), hi = min(n, lo + chunk);
            int c = 0;
            for (int i = lo; i < hi; i++) c += (int)(desc_key(lg[i]) >> (32 - SP_HBITS)) <= bstar;
            int pos = block_excl_scan(c, warp_tot, tid);
            for (int i = lo; i < hi; i++) {
                const uint32_t k = desc_key(lg[i]);
                if ((int)(k >> (32 - SP_HBITS)) <= bstar) { A.keys1[pos] = k; A.idx1[pos] = i; pos++; }
            }
            __syncthreads();
            radix_order(A.keys1, A.idx1, A.keys0, A.idx0, n_ord, cnt, warp_tot, tid);
        } else {
            n_ord = n;
            for (int i = tid; i < n; i += SP_T) { A.keys0[i] = desc_key(lg[i]); A.idx0[i] = i; }
            __syncthreads();
            radix_order(A.keys0, A.idx0, A.keys1, A.idx1, n, cnt, warp_tot, tid);
        }
        const int32_t *ord = attempt == 0 ? A.idx1 : A.idx0;
        const int limit = topp ? n_ord : min(n_ord, A.top_k);

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
        const int found = topp ? s_found : limit - 1;
        const float cum_total = s_cum;
        __syncthreads();
        if (found < 0) {
            if (attempt == 0 && n_ord < n) continue;   // (uniform) the nucleus is larger than the candidate set: order everything
            // the probabilities never reach top_p (rounding): the reference returns the most likely token (:397)
            if (tid == 0) *A.token_out = ord[0];
            return;
        }
        const float r = __fmul_rn(A.u, cum_total);   // rng.Float32() * cumsum (:339, :388)
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
        return;
    }
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
