// nl_kernels.cuh — decode-path kernels (sm_100a).  Each kernel cites the reference code it replaces.
#pragma once
#include "nl_common.cuh"
#include "nl_gemv_kernel.cuh"

namespace nl {

// =====================================================================================================
// Repack: raw GGUF blocks -> planar layout (see DevMat).  One thread per block.
// =====================================================================================================
static __global__ void repack_q4_0_kernel(const uint8_t *__restrict__ raw, uint4 *__restrict__ qs, __half *__restrict__ d, int64_t nblocks) {
    int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblocks) return;
    const uint16_t *p = reinterpret_cast<const uint16_t *>(raw + b * 18);  // blocks are 2-B aligned
    uint16_t h[9];
#pragma unroll
    for (int i = 0; i < 9; i++) h[i] = p[i];
    reinterpret_cast<uint16_t *>(d)[b] = h[0];
    uint4 q;
    q.x = h[1] | ((uint32_t)h[2] << 16); q.y = h[3] | ((uint32_t)h[4] << 16);
    q.z = h[5] | ((uint32_t)h[6] << 16); q.w = h[7] | ((uint32_t)h[8] << 16);
    qs[b] = q;
}
static __global__ void repack_q8_0_kernel(const uint8_t *__restrict__ raw, uint4 *__restrict__ qs, __half *__restrict__ d, int64_t nblocks) {
    int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblocks) return;
    const uint16_t *p = reinterpret_cast<const uint16_t *>(raw + b * 34);
    uint16_t h[17];
#pragma unroll
    for (int i = 0; i < 17; i++) h[i] = p[i];
    reinterpret_cast<uint16_t *>(d)[b] = h[0];
    uint4 q0, q1;
    q0.x = h[1] | ((uint32_t)h[2] << 16); q0.y = h[3] | ((uint32_t)h[4] << 16);
    q0.z = h[5] | ((uint32_t)h[6] << 16); q0.w = h[7] | ((uint32_t)h[8] << 16);
    q1.x = h[9] | ((uint32_t)h[10] << 16); q1.y = h[11] | ((uint32_t)h[12] << 16);
    q1.z = h[13] | ((uint32_t)h[14] << 16); q1.w = h[15] | ((uint32_t)h[16] << 16);
    qs[2 * b] = q0;
    qs[2 * b + 1] = q1;
}

// =====================================================================================================
// Dequant: planar -> fp32.  Bit-exact restatement of DequantQ4_0Block / DequantQ8_0Block / half2float
// (go/quant.go:22-31, :103-108, go/gguf.go:601-636).  (int - 8) * d is a single fp32 multiply, as in Go.
// =====================================================================================================
static __global__ void dequant_q4_0_kernel(const uint4 *__restrict__ qs, const __half *__restrict__ d, float *__restrict__ out, int64_t nblocks) {
    int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblocks) return;
    float s = h2f_exact(d[b]);
    uint4 q = qs[b];
    uint32_t w[4] = {q.x, q.y, q.z, q.w};
    float *o = out + b * 32;
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int k = 0; k < 4; k++) {
            uint32_t byte = (w[i] >> (8 * k)) & 0xFF;
            o[4 * i + k] = (float)((int)(byte & 0x0F) - 8) * s;
            o[4 * i + k + 16] = (float)((int)(byte >> 4) - 8) * s;
        }
}
static __global__ void dequant_q8_0_kernel(const uint4 *__restrict__ qs, const __half *__restrict__ d, float *__restrict__ out, int64_t nblocks) {
    int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblocks) return;
    float s = h2f_exact(d[b]);
    float *o = out + b * 32;
#pragma unroll
    for (int h = 0; h < 2; h++) {
        uint4 q = qs[2 * b + h];
        uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int k = 0; k < 4; k++) o[16 * h + 4 * i + k] = (float)(int8_t)((w[i] >> (8 * k)) & 0xFF) * s;
    }
}
static __global__ void dequant_f16_kernel(const __half *__restrict__ src, float *__restrict__ out, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = h2f_exact(src[i]);
}
// raw-block types (the "next" row of SURVEY §8f): go/quant.go:405-420 (Q5_0), :296-323 (Q4_K), :174-208 (Q6_K)
__device__ __forceinline__ float h2f_raw(const uint8_t *p) { return h2f_exact((unsigned short)(p[0] | (p[1] << 8))); }
__device__ __forceinline__ void scale_min_k4(int j, const uint8_t *s, int &sc, int &m) {
    if (j < 4) { sc = s[j] & 63; m = s[j + 4] & 63; }
    else { sc = (s[j + 4] & 0x0F) | ((s[j - 4] >> 6) << 4); m = (s[j + 4] >> 4) | ((s[j] >> 6) << 4); }
}
// one element of a raw block; used by both the dequant kernel and the generic GEMV so the value is identical
__device__ __forceinline__ float raw_elem(int type, const uint8_t *blk, int e) {
    if (type == NL_Q5_0) {
        float d = h2f_raw(blk);
        uint32_t qh = blk[2] | (blk[3] << 8) | (blk[4] << 16) | ((uint32_t)blk[5] << 24);
        int j = e & 15;
        int q = (e < 16) ? ((blk[6 + j] & 0x0F) | (((qh >> j) & 1) << 4)) : ((blk[6 + j] >> 4) | (((qh >> (j + 16)) & 1) << 4));
        return (float)(q - 16) * d;
    } else if (type == NL_Q4_K) {
        float d = h2f_raw(blk), dmin = h2f_raw(blk + 2);
        int grp = e >> 6, l = e & 31, hi = (e >> 5) & 1;
        int sc, m;
        scale_min_k4(2 * grp + hi, blk + 4, sc, m);
        float dd = d * (float)sc, mm = dmin * (float)m;
        uint8_t q = blk[16 + grp * 32 + l];
        return __fsub_rn(__fmul_rn(dd, (float)(hi ? (q >> 4) : (q & 0x0F))), mm);  // d1*q - m1, two roundings like Go
    } else {  // NL_Q6_K
        const uint8_t *ql = blk, *qh = blk + 128, *scales = blk + 192;
        float d = h2f_raw(blk + 208);
        int n128 = e >> 7, r = e & 127, quarter = r >> 5, l = r & 31, is = l >> 4;
        const uint8_t *qlP = ql + n128 * 64, *qhP = qh + n128 * 32, *scP = scales + n128 * 8;
        int q;
        if (quarter == 0) q = (qlP[l] & 0x0F) | (((qhP[l] >> 0) & 3) << 4);
        else if (quarter == 1) q = (qlP[l + 32] & 0x0F) | (((qhP[l] >> 2) & 3) << 4);
        else if (quarter == 2) q = (qlP[l] >> 4) | (((qhP[l] >> 4) & 3) << 4);
        else q = (qlP[l + 32] >> 4) | (((qhP[l] >> 6) & 3) << 4);
        return __fmul_rn(__fmul_rn(d, (float)(int8_t)scP[is + 2 * quarter]), (float)(q - 32));
    }
}
static __global__ void dequant_raw_kernel(int type, const uint8_t *__restrict__ raw, float *__restrict__ out, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int be = blk_elems(type), bb = blk_bytes(type);
    out[i] = raw_elem(type, raw + (i / be) * bb, (int)(i % be));
}

// generic (slow) GEMV for the raw-block types Q5_0 / Q4_K / Q6_K: one warp per row
template <int NB>
static __global__ void gemv_raw_kernel(int type, const uint8_t *__restrict__ raw, const float *__restrict__ x, int x_stride, float *__restrict__ out,
                                int out_stride, const float *__restrict__ bias, int rows, int cols, int epi) {
    int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    int be = blk_elems(type), bb = blk_bytes(type);
    const uint8_t *rp = raw + (int64_t)row * (cols / be) * bb;
    float acc[NB];
#pragma unroll
    for (int b = 0; b < NB; b++) acc[b] = 0.f;
    for (int k = lane; k < cols; k += 32) {
        float w = raw_elem(type, rp + (k / be) * bb, k % be);
#pragma unroll
        for (int b = 0; b < NB; b++) acc[b] = fmaf(w, x[(int64_t)b * x_stride + k], acc[b]);
    }
#pragma unroll
    for (int b = 0; b < NB; b++) {
        float v = warp_sum(acc[b]);
        if (lane == 0) {
            if (bias) v += bias[row];
            float *o = out + (int64_t)b * out_stride + row;
            if (epi == EPI_RESID) *o += v; else *o = v;
        }
    }
}

// =====================================================================================================
// Embedding row lookup (+ optional gamma row add): embedLookupInto, go/model.go:389-446, :503-507
// =====================================================================================================
static __global__ void embed_kernel(DevMat e, const int32_t *__restrict__ tokens, const float *__restrict__ gamma, const int32_t *__restrict__ gamma_map,
                             float *__restrict__ x, int dim) {
    const int b = blockIdx.y;
    const int tok = tokens[b];
    float *o = x + (int64_t)b * dim;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < dim; i += gridDim.x * blockDim.x) {
        float v;
        int64_t idx = (int64_t)tok * dim + i;
        if (e.type == NL_Q4_0) {
            int64_t blk = idx >> 5; int el = (int)(idx & 31);
            uint8_t byte = e.qs[blk * 16 + (el & 15)];
            v = (float)((int)(el < 16 ? (byte & 0x0F) : (byte >> 4)) - 8) * h2f_exact(e.d[blk]);
        } else if (e.type == NL_Q8_0) {
            v = (float)(int8_t)e.qs[idx] * h2f_exact(e.d[idx >> 5]);
        } else if (e.type == NL_F16) {
            v = h2f_exact(reinterpret_cast<const __half *>(e.qs)[idx]);
        } else if (e.type == NL_F32) {
            v = reinterpret_cast<const float *>(e.qs)[idx];
        } else {
            int be = blk_elems(e.type), bb = blk_bytes(e.type);
            v = raw_elem(e.type, e.qs + (idx / be) * bb, (int)(idx % be));
        }
        if (gamma_map) { int g = gamma_map[tok]; if (g >= 0) v += gamma[(int64_t)g * dim + i]; }
        if (__float_as_uint(v) == 0xFFFFFFFFu) v = __uint_as_float(0x7FFFFFFFu);   // (one NaN pattern is the sentinel of the polled vectors, nl_tile.cu)
        o[i] = v;
    }
}

// =====================================================================================================
// RMSNorm with weight: RMSNormInto / RMSNorm, go/quant.go:570-607.  float64 sum of squares, fp32 x*inv*w.
// one CTA per sequence.
// =====================================================================================================
static __global__ void __launch_bounds__(1024) rmsnorm_kernel(const float *__restrict__ x, const float *__restrict__ w, float *__restrict__ out, int n, float eps) {
    const float *xi = x + (int64_t)blockIdx.x * n;
    float *oi = out + (int64_t)blockIdx.x * n;
    double ss = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) { double v = (double)xi[i]; ss += v * v; }
    __shared__ double red[32];
    __shared__ float inv_s;
    ss = warp_sum_d(ss);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
        v = warp_sum_d(v);
        if (threadIdx.x == 0) inv_s = (float)(1.0 / sqrt(v / (double)n + (double)eps));
    }
    __syncthreads();
    const float inv = inv_s;
    for (int i = threadIdx.x; i < n; i += blockDim.x) oi[i] = xi[i] * inv * w[i];
}

// =====================================================================================================
// Decode attention for one (head, sequence): RoPE on q and on the new k (go/model.go:449-477, :530-539),
// optional bare QK-norm (:542-549, go/quant.go:584-594), KV-cache write (:552-554), scores, softmax
// (go/quant.go:610-626) and the weighted V sum (:557-587).  kvh = h / (H/KV).
// Every CTA ropes the new k/v head itself (so there is no cross-CTA dependency on the cache row being written);
// only the first head of each GQA group stores it.  KV cache is fp32 [seq][layer][pos][kv_dim] like the reference.
// =====================================================================================================
struct AttnArgs {
    const float *q, *k, *v;      // [B][H*hd], [B][kvd], [B][kvd]  (fresh projections of this token)
    float *kcache, *vcache;      // this layer's slab of sequence 0; + b * seq_stride for sequence b
    int64_t seq_stride;          // floats between sequences
    const float *cos_t, *sin_t;  // [seq_len][hd/2]
    const int32_t *pos;          // [B]
    float *out;                  // [B][H*hd]
    int n_heads, n_kv_heads, seq_len, qk_norm, conj;
    float eps, scale;
};

template <int HD>
static __global__ void __launch_bounds__(128) attn_decode_kernel(const AttnArgs a) {
    constexpr int HALF = HD / 2;
    const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    const int group = a.n_heads / a.n_kv_heads, kvh = h / group, kvd = a.n_kv_heads * HD;
    const int pos = a.pos[b];
    extern __shared__ float sc[];  // [seq_len] scores
    __shared__ float sq[HD], sk[HD], sv[HD], red[8], part[2][HD];
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // let the next (PDL) GEMV start streaming its weights

    // --- load + RoPE: threads [0,HALF) rotate q, [HALF,2*HALF) rotate k, next HD threads copy v
    const float *cs = a.cos_t + (int64_t)pos * HALF, *sn = a.sin_t + (int64_t)pos * HALF;
    if (tid < 2 * HALF) {
        const bool isk = tid >= HALF;
        const int i = isk ? tid - HALF : tid;
        const float *src = isk ? a.k + (int64_t)b * kvd + kvh * HD : a.q + ((int64_t)b * a.n_heads + h) * HD;
        float x0 = src[i], x1 = src[i + HALF], c = cs[i], s = sn[i];
        float r0, r1;
        if (!a.conj) { r0 = x0 * c - x1 * s; r1 = x0 * s + x1 * c; }
        else { r0 = x0 * c + x1 * s; r1 = -x0 * s + x1 * c; }
        float *dst = isk ? sk : sq;
        dst[i] = r0; dst[i + HALF] = r1;
    }
    for (int i = tid; i < HD; i += 128) sv[i] = a.v[(int64_t)b * kvd + kvh * HD + i];
    __syncthreads();
    if (a.qk_norm) {  // RMSNormBare on q head and k head: warp 0 -> q, warp 1 -> k
        if (tid < 64) {
            float *vec = tid < 32 ? sq : sk;
            const int l = tid & 31;
            double ss = 0.0;
            for (int i = l; i < HD; i += 32) ss += (double)vec[i] * (double)vec[i];
            ss = warp_sum_d(ss);
            const float inv = (float)(1.0 / sqrt(ss / (double)HD + (double)a.eps));
            for (int i = l; i < HD; i += 32) vec[i] *= inv;
        }
        __syncthreads();
    }
    float *kc = a.kcache + (int64_t)b * a.seq_stride, *vc = a.vcache + (int64_t)b * a.seq_stride;
    if (h % group == 0)
        for (int i = tid; i < HD; i += 128) {
            kc[(int64_t)pos * kvd + kvh * HD + i] = sk[i];
            vc[(int64_t)pos * kvd + kvh * HD + i] = sv[i];
        }

    // --- scores: 8 lanes per cached position, each lane HD/8 contiguous floats
    constexpr int PER = HD / 8;
    const int sub = tid & 7, tg = tid >> 3;  // 16 positions per pass
    float qreg[PER];
#pragma unroll
    for (int i = 0; i < PER; i++) qreg[i] = sq[sub * PER + i];
    for (int t0 = 0; t0 <= pos; t0 += 16) {
        const int t = t0 + tg;
        float dot = 0.f;
        if (t <= pos) {
            if (t < pos) {
                const float4 *kp = reinterpret_cast<const float4 *>(kc + (int64_t)t * kvd + kvh * HD + sub * PER);
#pragma unroll
                for (int i = 0; i < PER / 4; i++) {
                    float4 kk = kp[i];
                    dot = fmaf(qreg[4 * i], kk.x, dot); dot = fmaf(qreg[4 * i + 1], kk.y, dot);
                    dot = fmaf(qreg[4 * i + 2], kk.z, dot); dot = fmaf(qreg[4 * i + 3], kk.w, dot);
                }
            } else {
#pragma unroll
                for (int i = 0; i < PER; i++) dot = fmaf(qreg[i], sk[sub * PER + i], dot);
            }
        }
        dot += __shfl_xor_sync(0xffffffffu, dot, 1);
        dot += __shfl_xor_sync(0xffffffffu, dot, 2);
        dot += __shfl_xor_sync(0xffffffffu, dot, 4);
        if (sub == 0 && t <= pos) sc[t] = dot * a.scale;
    }
    __syncthreads();

    // --- softmax over sc[0..pos]
    const int n = pos + 1, lane = tid & 31, warp = tid >> 5;
    float mx = -INFINITY;
    for (int t = tid; t < n; t += 128) mx = fmaxf(mx, sc[t]);
    mx = warp_max(mx);
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
    float sum = 0.f;
    for (int t = tid; t < n; t += 128) { float e = expf(sc[t] - mx); sc[t] = e; sum += e; }
    sum = warp_sum(sum);
    if (lane == 0) red[4 + warp] = sum;
    __syncthreads();
    const float inv = 1.0f / (red[4] + red[5] + red[6] + red[7]);

    // --- weighted V sum: thread = (dim d, position parity)
    {
        constexpr int NPAR = 128 / HD;           // HD=64: 2 position parities; HD=128: 1
        const int d = tid % HD, par = tid / HD;
        float acc = 0.f;
        for (int t = par; t < n; t += NPAR) {
            float vv = (t < pos) ? vc[(int64_t)t * kvd + kvh * HD + d] : sv[d];
            acc = fmaf(sc[t] * inv, vv, acc);
        }
        part[par][d] = acc;
    }
    __syncthreads();
    if (tid < HD) {
        float o = part[0][tid];
        if (128 / HD >= 2) o += part[1][tid];
        a.out[((int64_t)b * a.n_heads + h) * HD + tid] = o;
    }
}

// =====================================================================================================
// argmax with the reference's tie rule (first maximum, strict '>': go/main.go:400-408) + device-side feedback
// for the greedy loop: writes the chosen token as the next input token and advances the position.
// =====================================================================================================
struct StepState {
    int32_t *token;      // [B] next input token
    int32_t *pos;        // [B] position of the next forward
    int32_t *gen;        // [B][gen_cap] generated tokens
    int32_t *gen_count;  // [B]
    int gen_cap;
};

static __global__ void __launch_bounds__(1024) argmax_advance_kernel(const float *__restrict__ logits, int vocab, StepState st) {
    const int b = blockIdx.x;
    const float *l = logits + (int64_t)b * vocab;
    float best = -INFINITY; int bi = 0x7fffffff;
    for (int i = threadIdx.x; i < vocab; i += blockDim.x) {
        float v = l[i];
        if (v > best || (v == best && i < bi)) { best = v; bi = i; }
    }
    __shared__ float sv[32]; __shared__ int si[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, best, o); int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = best; si[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x < 32) {
        best = threadIdx.x < (blockDim.x >> 5) ? sv[threadIdx.x] : -INFINITY;
        bi = threadIdx.x < (blockDim.x >> 5) ? si[threadIdx.x] : 0x7fffffff;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            float ov = __shfl_xor_sync(0xffffffffu, best, o); int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if (threadIdx.x == 0) {
            if (bi == 0x7fffffff) bi = 0;  // all -inf / NaN: Go's loop would keep index 0
            const int g = st.gen_count[b];
            if (g < st.gen_cap) st.gen[(int64_t)b * st.gen_cap + g] = bi;
            st.token[b] = bi;
            st.pos[b] += 1;
            st.gen_count[b] = g + 1;
        }
    }
}
// same step when the persistent decode kernel has already left one (maximum, first index) pair per CTA (nl_tile.cu): n pairs, one warp
static __global__ void __launch_bounds__(32) argmax_pairs_advance_kernel(const float2 *__restrict__ pairs, int n, StepState st) {
    float best = -INFINITY; int bi = 0x7fffffff;
    for (int i = threadIdx.x; i < n; i += 32) {
        const float2 pr = pairs[i];
        const float v = pr.x; const int vi = __float_as_int(pr.y);
        if (v > best || (v == best && vi < bi)) { best = v; bi = vi; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, best, o); int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (threadIdx.x == 0) {
        if (bi == 0x7fffffff) bi = 0;  // all -inf / NaN: Go's loop would keep index 0
        const int g = st.gen_count[0];
        if (g < st.gen_cap) st.gen[g] = bi;
        st.token[0] = bi;
        st.pos[0] += 1;
        st.gen_count[0] = g + 1;
    }
}
// prefill feed: token <- prompt[i], pos <- i for sequence 0
static __global__ void feed_prompt_kernel(const int32_t *__restrict__ prompt, int32_t *cursor, int32_t pos0, int32_t *token, int32_t *pos) {
    const int i = *cursor;
    token[0] = prompt[i];
    pos[0] = pos0 + i;
    *cursor = i + 1;
}

}  // namespace nl
