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
                                                                   __nv_bfloat16 *__restrict__ lo, int n, float eps) {
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
        *reinterpret_cast<uint32_t *>(hi + (size_t)blockIdx.x * n + i) = h;
        *reinterpret_cast<uint32_t *>(lo + (size_t)blockIdx.x * n + i) = l;
    }
}

// planes = split_bf16(SiLU(gate) * up)                           go/model.go:604-606
static __global__ void swiglu_split_kernel(const float *__restrict__ gate, const float *__restrict__ up, __nv_bfloat16 *__restrict__ hi,
                                           __nv_bfloat16 *__restrict__ lo, int64_t n) {
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
    if (i + 1 < n) {
        uint32_t h, l;
        split2(silu_f(gate[i]) * up[i], silu_f(gate[i + 1]) * up[i + 1], h, l);
        *reinterpret_cast<uint32_t *>(hi + i) = h;
        *reinterpret_cast<uint32_t *>(lo + i) = l;
    }
}

struct PrefillAttn {
    float *qkv;                  // [T][ld]: q at col 0, k at col qdim, v at col qdim + kvd
    int ld, T, pos0;
    float *kcache, *vcache;      // this layer's slab [S][kvd]
    const float *cos_t, *sin_t;  // [S][32]
    __nv_bfloat16 *out_hi, *out_lo;  // [T][qdim]
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
        const size_t base = (size_t)t * qdim + h * HD;
        a.out_hi[base + lane] = h0; a.out_hi[base + lane + 32] = h1;
        a.out_lo[base + lane] = __float2bfloat16_rn(o0 - __bfloat162float(h0));
        a.out_lo[base + lane + 32] = __float2bfloat16_rn(o1 - __bfloat162float(h1));
    }
}

}  // namespace nl
