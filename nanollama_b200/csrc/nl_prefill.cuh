// nl_prefill.cuh — the non-GEMM kernels of the one-pass prefill (T prompt tokens at once): row RMSNorm -> bf16 planes,
// RoPE + QK-norm + KV-cache write for T positions, causal attention over the fp32 KV cache, SiLU*up -> bf16 planes.
// Same arithmetic as the per-token path (go/model.go:517-606), T rows at a time; the matmuls are nl_gemm.cuh.
#pragma once
#include <cuda_bf16.h>

#include "nl_common.cuh"
#include "nl_gemm.cuh"

namespace nl {

// out planes [T][n] = split_bf16(x[t] * inv_rms(x[t]) * w)      RMSNormInto, go/quant.go:597-607
static __global__ void __launch_bounds__(256) rmsnorm_split_kernel(const float *__restrict__ x, const float *__restrict__ w, __nv_bfloat16 *__restrict__ hi,
                                                                   __nv_bfloat16 *__restrict__ lo, int n, float eps, int tiled) {
    const float *xi = x + (size_t)blockIdx.x * n;
    double ss = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) { const double v = (double)xi[i]; ss += v * v; }
    __shared__ double red[8];
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
    for (int i = threadIdx.x * 2; i < n; i += blockDim.x * 2) {
        uint32_t h, l;
        split2(xi[i] * inv * w[i], xi[i + 1] * inv * w[i + 1], h, l);
        const size_t o = plane_index((int)blockIdx.x, i, n, tiled);
        *reinterpret_cast<uint32_t *>(hi + o) = h;
        *reinterpret_cast<uint32_t *>(lo + o) = l;
    }
}

// planes = split_bf16(SiLU(gate) * up)                           go/model.go:604-606
static __global__ void swiglu_split_kernel(const float *__restrict__ gate, const float *__restrict__ up, __nv_bfloat16 *__restrict__ hi,
                                           __nv_bfloat16 *__restrict__ lo, int64_t n, int K, int tiled) {
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
    if (i + 1 < n) {
        uint32_t h, l;
        split2(silu_f(gate[i]) * up[i], silu_f(gate[i + 1]) * up[i + 1], h, l);
        const size_t o = tiled ? plane_index((int)(i / K), (int)(i % K), K, 1) : (size_t)i;   // (K is even: a pair never straddles two rows)
        *reinterpret_cast<uint32_t *>(hi + o) = h;
        *reinterpret_cast<uint32_t *>(lo + o) = l;
    }
}

struct PrefillAttn {
    float *qkv;                  // [T][ld]: q at col 0, k at col qdim, v at col qdim + kvd
    int ld, T, pos0;
    float *kcache, *vcache;      // this layer's slab [S][kvd]
    const float *cos_t, *sin_t;  // [S][32]
    __nv_bfloat16 *out_hi, *out_lo;  // [T][qdim] (tiled: plane_index order)
    int tiled;
    int n_heads, n_kv_heads, qk_norm, conj;
    float eps, scale;
};

// one CTA per prompt position: RoPE on every q and k head (go/model.go:449-477, :530-539), bare QK-norm (:542-549),
// K and V rows into the cache (:552-554).  One warp per head, lane i owns the pair (i, i+32).  head_dim = 64.
static __global__ void __launch_bounds__(256) rope_kv_kernel(const PrefillAttn a) {
    const int t = blockIdx.x, pos = a.pos0 + t, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qdim = a.n_heads * 64, kvd = a.n_kv_heads * 64;
    float *row = a.qkv + (size_t)t * a.ld;
    const float c = a.cos_t[(size_t)pos * 32 + lane], s = a.sin_t[(size_t)pos * 32 + lane];
    for (int h = warp; h < a.n_heads + a.n_kv_heads; h += 8) {
        float *v = h < a.n_heads ? row + h * 64 : row + qdim + (h - a.n_heads) * 64;
        const float x0 = v[lane], x1 = v[lane + 32];
        float r0, r1;
        if (!a.conj) { r0 = x0 * c - x1 * s; r1 = x0 * s + x1 * c; }
        else { r0 = x0 * c + x1 * s; r1 = -x0 * s + x1 * c; }
        if (a.qk_norm) {
            double ss = (double)r0 * (double)r0 + (double)r1 * (double)r1;
            ss = warp_sum_d(ss);
            const float inv = (float)(1.0 / sqrt(ss / 64.0 + (double)a.eps));
            r0 *= inv; r1 *= inv;
        }
        v[lane] = r0; v[lane + 32] = r1;
        if (h >= a.n_heads) {
            float *kc = a.kcache + (size_t)pos * kvd + (h - a.n_heads) * 64;
            kc[lane] = r0; kc[lane + 32] = r1;
        }
    }
    for (int i = threadIdx.x; i < kvd; i += 256) a.vcache[(size_t)pos * kvd + i] = row[qdim + kvd + i];
}

// causal attention for 32 consecutive prompt positions of one head against the fp32 KV cache (go/model.go:557-587);
// 8 warps x 4 queries, key tiles of 32 staged in shared memory, running softmax, output as bf16 planes for the O GEMM.
static __global__ void __launch_bounds__(256) attn_prefill_kernel(const PrefillAttn a) {
    constexpr int HD = 64, QT = 32, KT = 32, PAD = HD + 1;
    const int h = blockIdx.x, q0 = blockIdx.y * QT, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int group = a.n_heads / a.n_kv_heads, kvh = h / group, kvd = a.n_kv_heads * HD, qdim = a.n_heads * HD;
    __shared__ float sq[QT][HD];
    __shared__ float sk[KT][PAD], sv[KT][PAD];
    __shared__ float sp[8][KT];
    for (int i = threadIdx.x; i < QT * HD; i += 256) {
        const int qi = i / HD, d = i % HD;
        sq[qi][d] = (q0 + qi < a.T) ? a.qkv[(size_t)(q0 + qi) * a.ld + h * HD + d] : 0.f;
    }
    float m_run[4], l_run[4], acc0[4], acc1[4];
#pragma unroll
    for (int j = 0; j < 4; j++) { m_run[j] = -INFINITY; l_run[j] = 0.f; acc0[j] = 0.f; acc1[j] = 0.f; }
    const int last_q = min(q0 + QT, a.T) - 1;
    const int n_keys = a.pos0 + last_q + 1;       // keys visible to the last query of this tile
    for (int k0 = 0; k0 < n_keys; k0 += KT) {
        __syncthreads();
        for (int i = threadIdx.x; i < KT * HD / 4; i += 256) {
            const int kj = i / (HD / 4), d4 = (i % (HD / 4)) * 4;
            float4 kk = make_float4(0, 0, 0, 0), vv = make_float4(0, 0, 0, 0);
            if (k0 + kj < n_keys) {
                kk = *reinterpret_cast<const float4 *>(a.kcache + (size_t)(k0 + kj) * kvd + kvh * HD + d4);
                vv = *reinterpret_cast<const float4 *>(a.vcache + (size_t)(k0 + kj) * kvd + kvh * HD + d4);
            }
            sk[kj][d4] = kk.x; sk[kj][d4 + 1] = kk.y; sk[kj][d4 + 2] = kk.z; sk[kj][d4 + 3] = kk.w;
            sv[kj][d4] = vv.x; sv[kj][d4 + 1] = vv.y; sv[kj][d4 + 2] = vv.z; sv[kj][d4 + 3] = vv.w;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int qi = warp * 4 + j, t = q0 + qi;
            if (t >= a.T) continue;                       // warp-uniform
            const int p = a.pos0 + t, key = k0 + lane;
            float sc = -INFINITY;
            if (key <= p) {
                float dot = 0.f;
#pragma unroll 16
                for (int d = 0; d < HD; d++) dot = fmaf(sq[qi][d], sk[lane][d], dot);
                sc = dot * a.scale;
            }
            const float mx = warp_max(sc);
            if (mx == -INFINITY) continue;                // whole tile is in the future of this query (warp-uniform)
            const float m_new = fmaxf(m_run[j], mx);
            const float e = (sc == -INFINITY) ? 0.f : expf(sc - m_new);
            const float corr = (m_run[j] == -INFINITY) ? 0.f : expf(m_run[j] - m_new);
            l_run[j] = l_run[j] * corr + warp_sum(e);
            m_run[j] = m_new;
            sp[warp][lane] = e;
            __syncwarp();
            float a0 = acc0[j] * corr, a1 = acc1[j] * corr;
#pragma unroll 8
            for (int kj = 0; kj < KT; kj++) {
                const float pj = sp[warp][kj];
                a0 = fmaf(pj, sv[kj][lane], a0);
                a1 = fmaf(pj, sv[kj][lane + 32], a1);
            }
            acc0[j] = a0; acc1[j] = a1;
            __syncwarp();
        }
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int t = q0 + warp * 4 + j;
        if (t >= a.T) continue;
        const float inv = 1.0f / l_run[j];
        const float o0 = acc0[j] * inv, o1 = acc1[j] * inv;
        const __nv_bfloat16 h0 = __float2bfloat16_rn(o0), h1 = __float2bfloat16_rn(o1);
        const size_t b0 = plane_index(t, h * HD + lane, qdim, a.tiled), b1 = plane_index(t, h * HD + lane + 32, qdim, a.tiled);
        a.out_hi[b0] = h0; a.out_hi[b1] = h1;
        a.out_lo[b0] = __float2bfloat16_rn(o0 - __bfloat162float(h0));
        a.out_lo[b1] = __float2bfloat16_rn(o1 - __bfloat162float(h1));
    }
}

// ---- the same attention on the tensor cores (flash-attention form): 128 consecutive prompt positions of one head per CTA, 8 warps x
// 16 query rows, key / value tiles of 64 positions converted from the fp32 cache to bf16 hi | lo planes in shared memory.  Every
// product is issued as three bf16 MMAs (hi*hi + hi*lo + lo*hi, mma.sync.m16n8k16, fp32 accumulate: ~2^-16 relative, the same split
// as the GEMMs of nl_gemm.cuh), S = Q K^T and O += P V stay in registers, the softmax runs on the accumulator fragments (running
// maximum / sum per row, go/quant.go:610-626 in its streaming form).  Replaces attn_prefill_kernel (CUDA cores, one shared-memory
// load per multiply-add), which was 44 % of the prefill (profiles/r01_launches_prefill_goldie.md).
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
constexpr int PA_QT = 128, PA_KT = 64, PA_LD = 72;   // queries per CTA, keys per tile, padded row length (bf16) of the shared tiles

static __global__ void __launch_bounds__(256) attn_prefill_tc_kernel(const PrefillAttn a) {
    constexpr int HD = 64;
    __shared__ __align__(16) __nv_bfloat16 sKh[PA_KT][PA_LD], sKl[PA_KT][PA_LD], sVh[PA_KT][PA_LD], sVl[PA_KT][PA_LD];
    const int h = blockIdx.x, q0 = blockIdx.y * PA_QT, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int group = a.n_heads / a.n_kv_heads, kvh = h / group, kvd = a.n_kv_heads * HD, qdim = a.n_heads * HD;
    const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;   // my two query rows (prompt indices)
    // Q fragments (A operand, 16 x 64 per warp): hi | lo planes, four k-steps of 16 dims
    uint32_t qh[4][4], ql[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ks++) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int row = (j & 1) ? r1 : r0, col = 16 * ks + 2 * t + ((j & 2) ? 8 : 0);
            float2 v = make_float2(0.f, 0.f);
            if (row < a.T) v = *reinterpret_cast<const float2 *>(a.qkv + (size_t)row * a.ld + h * HD + col);
            split2(v.x, v.y, qh[ks][j], ql[ks][j]);
        }
    }
    float o[8][4];
#pragma unroll
    for (int d = 0; d < 8; d++) { o[d][0] = 0.f; o[d][1] = 0.f; o[d][2] = 0.f; o[d][3] = 0.f; }
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;   // running maximum / (per-lane part of the) sum of rows r0, r1
    const int last_q = min(q0 + PA_QT, a.T) - 1;
    const int n_keys = a.pos0 + last_q + 1;                     // keys visible to the last query of this CTA
    const int p0 = a.pos0 + r0, p1 = a.pos0 + r1;               // my rows' positions: keys <= p are visible
    const uint32_t kh_u = (uint32_t)__cvta_generic_to_shared(&sKh[0][0]), kl_u = (uint32_t)__cvta_generic_to_shared(&sKl[0][0]);
    const uint32_t vh_u = (uint32_t)__cvta_generic_to_shared(&sVh[0][0]), vl_u = (uint32_t)__cvta_generic_to_shared(&sVl[0][0]);
    for (int k0 = 0; k0 < n_keys; k0 += PA_KT) {
        __syncthreads();   // the previous tile has been consumed
        for (int i = tid; i < PA_KT * HD / 4; i += 256) {        // fp32 cache rows -> bf16 hi | lo planes
            const int kj = i >> 4, d4 = (i & 15) * 4;
            float4 kk = make_float4(0.f, 0.f, 0.f, 0.f), vv = kk;
            if (k0 + kj < n_keys) {
                kk = *reinterpret_cast<const float4 *>(a.kcache + (size_t)(k0 + kj) * kvd + kvh * HD + d4);
                vv = *reinterpret_cast<const float4 *>(a.vcache + (size_t)(k0 + kj) * kvd + kvh * HD + d4);
            }
            uint2 hh, ll;
            split2(kk.x, kk.y, hh.x, ll.x); split2(kk.z, kk.w, hh.y, ll.y);
            *reinterpret_cast<uint2 *>(&sKh[kj][d4]) = hh; *reinterpret_cast<uint2 *>(&sKl[kj][d4]) = ll;
            split2(vv.x, vv.y, hh.x, ll.x); split2(vv.z, vv.w, hh.y, ll.y);
            *reinterpret_cast<uint2 *>(&sVh[kj][d4]) = hh; *reinterpret_cast<uint2 *>(&sVl[kj][d4]) = ll;
        }
        __syncthreads();
        if (q0 + warp * 16 >= a.T || k0 > a.pos0 + min(q0 + warp * 16 + 15, a.T - 1)) continue;   // (warp-uniform) nothing visible to this warp's rows
        // ---- S = Q K^T (16 x 64 per warp)
        float sc[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; nt++) {
            sc[nt][0] = 0.f; sc[nt][1] = 0.f; sc[nt][2] = 0.f; sc[nt][3] = 0.f;
#pragma unroll
            for (int p2 = 0; p2 < 2; p2++) {
                const uint32_t off = (uint32_t)(((nt * 8 + (lane & 7)) * PA_LD + (4 * p2 + (lane >> 3)) * 8) * 2);
                uint32_t bh[4], bl[4];
                ldsm_x4(kh_u + off, bh); ldsm_x4(kl_u + off, bl);
                mma_bf16(sc[nt], qh[2 * p2], bh[0], bh[1]); mma_bf16(sc[nt], qh[2 * p2], bl[0], bl[1]); mma_bf16(sc[nt], ql[2 * p2], bh[0], bh[1]);
                mma_bf16(sc[nt], qh[2 * p2 + 1], bh[2], bh[3]); mma_bf16(sc[nt], qh[2 * p2 + 1], bl[2], bl[3]); mma_bf16(sc[nt], ql[2 * p2 + 1], bh[2], bh[3]);
            }
        }
        // ---- scale, causal mask, running softmax (rows r0 / r1: elements 0,1 / 2,3 of every fragment)
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 8; nt++) {
            const int key = k0 + nt * 8 + 2 * t;
            sc[nt][0] = key <= p0 ? sc[nt][0] * a.scale : -INFINITY; sc[nt][1] = key + 1 <= p0 ? sc[nt][1] * a.scale : -INFINITY;
            sc[nt][2] = key <= p1 ? sc[nt][2] * a.scale : -INFINITY; sc[nt][3] = key + 1 <= p1 ? sc[nt][3] * a.scale : -INFINITY;
            mx0 = fmaxf(mx0, fmaxf(sc[nt][0], sc[nt][1])); mx1 = fmaxf(mx1, fmaxf(sc[nt][2], sc[nt][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float n0 = fmaxf(m0, mx0), n1 = fmaxf(m1, mx1);
        // (a row whose keys are all in the future so far keeps m = -inf: use 0 as the reference point so that exp() sees -inf, not nan)
        const float b0 = n0 == -INFINITY ? 0.f : n0, b1 = n1 == -INFINITY ? 0.f : n1;
        const float c0 = expf(m0 - b0), c1 = expf(m1 - b1);
        m0 = n0; m1 = n1;
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int nt = 0; nt < 8; nt++) {
            sc[nt][0] = expf(sc[nt][0] - b0); sc[nt][1] = expf(sc[nt][1] - b0); sc[nt][2] = expf(sc[nt][2] - b1); sc[nt][3] = expf(sc[nt][3] - b1);
            s0 += sc[nt][0] + sc[nt][1]; s1 += sc[nt][2] + sc[nt][3];
        }
        l0 = l0 * c0 + s0; l1 = l1 * c1 + s1;
#pragma unroll
        for (int d = 0; d < 8; d++) { o[d][0] *= c0; o[d][1] *= c0; o[d][2] *= c1; o[d][3] *= c1; }
        // ---- O += P V: the score fragments of key tiles 2kk, 2kk+1 ARE the A operand of key step kk
#pragma unroll
        for (int kk = 0; kk < 4; kk++) {
            uint32_t ph[4], pl[4];
            split2(sc[2 * kk][0], sc[2 * kk][1], ph[0], pl[0]); split2(sc[2 * kk][2], sc[2 * kk][3], ph[1], pl[1]);
            split2(sc[2 * kk + 1][0], sc[2 * kk + 1][1], ph[2], pl[2]); split2(sc[2 * kk + 1][2], sc[2 * kk + 1][3], ph[3], pl[3]);
#pragma unroll
            for (int d2 = 0; d2 < 4; d2++) {
                const int j = lane >> 3;
                const uint32_t off = (uint32_t)(((16 * kk + (j & 1) * 8 + (lane & 7)) * PA_LD + (2 * d2 + (j >> 1)) * 8) * 2);
                uint32_t vh[4], vl[4];
                ldsm_x4_t(vh_u + off, vh); ldsm_x4_t(vl_u + off, vl);
                mma_bf16(o[2 * d2], ph, vh[0], vh[1]); mma_bf16(o[2 * d2], ph, vl[0], vl[1]); mma_bf16(o[2 * d2], pl, vh[0], vh[1]);
                mma_bf16(o[2 * d2 + 1], ph, vh[2], vh[3]); mma_bf16(o[2 * d2 + 1], ph, vl[2], vl[3]); mma_bf16(o[2 * d2 + 1], pl, vh[2], vh[3]);
            }
        }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.0f / l0, i1 = 1.0f / l1;
#pragma unroll
    for (int d = 0; d < 8; d++) {
        const int col = h * HD + d * 8 + 2 * t;
        uint32_t hh, ll;
        if (r0 < a.T) {
            split2(o[d][0] * i0, o[d][1] * i0, hh, ll);
            const size_t o0 = plane_index(r0, col, qdim, a.tiled);
            *reinterpret_cast<uint32_t *>(a.out_hi + o0) = hh; *reinterpret_cast<uint32_t *>(a.out_lo + o0) = ll;
        }
        if (r1 < a.T) {
            split2(o[d][2] * i1, o[d][3] * i1, hh, ll);
            const size_t o1 = plane_index(r1, col, qdim, a.tiled);
            *reinterpret_cast<uint32_t *>(a.out_hi + o1) = hh; *reinterpret_cast<uint32_t *>(a.out_lo + o1) = ll;
        }
    }
}

}  // namespace nl

