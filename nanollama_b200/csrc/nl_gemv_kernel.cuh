// nl_gemv_kernel.cuh — the dequant-fused GEMV kernel template (sm_100a).
#pragma once
#include "nl_common.cuh"

namespace nl {

// =====================================================================================================
// Dequant-fused GEMV (decode):  out[b][row] = sum_k W[row][k] * x[b][k]    (matmulDispatch, go/model.go:361-386;
// MatMulQ4_0 / Q8_0 / F16 / F32, go/quant.go:45-165, :490-563).  fp32 activations, fp32 accumulate.
//
// Work decomposition: a CTA owns R consecutive rows; its THREADS threads stride over the row's 16-byte "units"
// (Q4_0: 8 quant bytes = 16 elems; Q8_0: 16 int8; F16: 8 halves; F32: 4 floats).  A thread keeps the x slice of its
// unit in registers and reuses it for the R rows, so weights are streamed exactly once (128-/64-bit coalesced,
// L1::no_allocate) and x comes from L2.  Per-row partials are combined with warp shuffles + one smem pass.
// Up to 3 independent segments (e.g. Wq/Wk/Wv, or gate/up) share one launch.
// =====================================================================================================
enum { EPI_STORE = 0, EPI_RESID = 1, EPI_SWIGLU = 2 };

struct GemvSeg {
    const uint8_t *qs;   // planar quants (or raw F16/F32 rows)
    const __half *d;     // block scales (quantized types)
    const uint8_t *qs2;  // EPI_SWIGLU: the "up" matrix paired with this "gate" matrix
    const __half *d2;
    const float *bias;   // optional (Qwen-style attention bias, go/model.go:480-487)
    float *out;          // [batch][out_stride]
    int rows;
    int out_stride;
    int cta_begin;       // first blockIdx.x of this segment
};
struct GemvArgs {
    GemvSeg seg[3];
    int nseg;
    const float *x;  // [batch][x_stride]
    int x_stride;
    int cols;
};

template <int TYPE> struct Unit;
template <> struct Unit<NL_Q4_0> { static constexpr int ELEMS = 16; };
template <> struct Unit<NL_Q8_0> { static constexpr int ELEMS = 16; };
template <> struct Unit<NL_F16> { static constexpr int ELEMS = 8; };
template <> struct Unit<NL_F32> { static constexpr int ELEMS = 4; };

// x slice for unit u: Q4_0 unit (blk, s) covers elems blk*32 + 8s + [0,8) (low nibbles) and blk*32 + 16 + 8s + [0,8) (high)
template <int TYPE>
__device__ __forceinline__ void load_x(const float *__restrict__ x, int u, float (&xr)[Unit<TYPE>::ELEMS]) {
    if constexpr (TYPE == NL_Q4_0) {
        const float *p = x + (u >> 1) * 32 + (u & 1) * 8;
        float4 a = *reinterpret_cast<const float4 *>(p), b = *reinterpret_cast<const float4 *>(p + 4);
        float4 c = *reinterpret_cast<const float4 *>(p + 16), e = *reinterpret_cast<const float4 *>(p + 20);
        xr[0] = a.x; xr[1] = a.y; xr[2] = a.z; xr[3] = a.w; xr[4] = b.x; xr[5] = b.y; xr[6] = b.z; xr[7] = b.w;
        xr[8] = c.x; xr[9] = c.y; xr[10] = c.z; xr[11] = c.w; xr[12] = e.x; xr[13] = e.y; xr[14] = e.z; xr[15] = e.w;
    } else {
        const float *p = x + u * Unit<TYPE>::ELEMS;
#pragma unroll
        for (int i = 0; i < Unit<TYPE>::ELEMS / 4; i++) {
            float4 a = *reinterpret_cast<const float4 *>(p + 4 * i);
            xr[4 * i] = a.x; xr[4 * i + 1] = a.y; xr[4 * i + 2] = a.z; xr[4 * i + 3] = a.w;
        }
    }
}

template <int TYPE> struct WUnit { uint4 q; unsigned short dbits; };

template <int TYPE>
__device__ __forceinline__ void load_w(const uint8_t *__restrict__ qs, const __half *__restrict__ d, int64_t row_units, int64_t row_blocks,
                                       int u, WUnit<TYPE> &w) {
    if constexpr (TYPE == NL_Q4_0) {
        uint2 v = ldg_stream_u2(qs + (row_units + u) * 8);
        w.q.x = v.x; w.q.y = v.y;
        w.dbits = ldg_stream_u16(d + row_blocks + (u >> 1));
    } else if constexpr (TYPE == NL_Q8_0) {
        w.q = ldg_stream_u4(qs + (row_units + u) * 16);
        w.dbits = ldg_stream_u16(d + row_blocks + (u >> 1));
    } else {
        w.q = ldg_stream_u4(qs + (row_units + u) * 16);
    }
}

template <int TYPE>
__device__ __forceinline__ float dot_unit(const WUnit<TYPE> &w, const float (&xr)[Unit<TYPE>::ELEMS]) {
    float acc = 0.f;
    if constexpr (TYPE == NL_Q4_0) {
        const uint32_t ws[2] = {w.q.x, w.q.y};
#pragma unroll
        for (int i = 0; i < 2; i++) {
            uint32_t lo = ws[i] & 0x0F0F0F0Fu, hi = (ws[i] >> 4) & 0x0F0F0F0Fu;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                acc = fmaf(u8_to_f32_magic(lo, k) - 8388616.0f, xr[4 * i + k], acc);      // (nibble - 8) * x
                acc = fmaf(u8_to_f32_magic(hi, k) - 8388616.0f, xr[8 + 4 * i + k], acc);
            }
        }
        return acc * __half2float(__ushort_as_half(w.dbits));
    } else if constexpr (TYPE == NL_Q8_0) {
        const uint32_t ws[4] = {w.q.x ^ 0x80808080u, w.q.y ^ 0x80808080u, w.q.z ^ 0x80808080u, w.q.w ^ 0x80808080u};
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int k = 0; k < 4; k++) acc = fmaf(u8_to_f32_magic(ws[i], k) - 8388736.0f, xr[4 * i + k], acc);  // int8 * x
        return acc * __half2float(__ushort_as_half(w.dbits));
    } else if constexpr (TYPE == NL_F16) {
        const uint32_t ws[4] = {w.q.x, w.q.y, w.q.z, w.q.w};
#pragma unroll
        for (int i = 0; i < 4; i++) {
            float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&ws[i]));
            acc = fmaf(f.x, xr[2 * i], acc);
            acc = fmaf(f.y, xr[2 * i + 1], acc);
        }
        return acc;
    } else {
        acc = fmaf(__uint_as_float(w.q.x), xr[0], acc); acc = fmaf(__uint_as_float(w.q.y), xr[1], acc);
        acc = fmaf(__uint_as_float(w.q.z), xr[2], acc); acc = fmaf(__uint_as_float(w.q.w), xr[3], acc);
        return acc;
    }
}


template <int TYPE, int R, int NB, int THREADS, int EPI>
__global__ void __launch_bounds__(THREADS) gemv_kernel(const GemvArgs a) {
    constexpr int NM = (EPI == EPI_SWIGLU) ? 2 : 1;
    constexpr int UE = Unit<TYPE>::ELEMS;
    // which segment does this CTA belong to?
    int s = 0;
    if (a.nseg > 1 && (int)blockIdx.x >= a.seg[1].cta_begin) s = 1;
    if (a.nseg > 2 && (int)blockIdx.x >= a.seg[2].cta_begin) s = 2;
    const GemvSeg &sg = a.seg[s];
    const int row0 = ((int)blockIdx.x - sg.cta_begin) * R;
    const int units = a.cols / UE;
    const int blocks_per_row = a.cols / 32;

    float acc[NM][R][NB];
#pragma unroll
    for (int m = 0; m < NM; m++)
#pragma unroll
        for (int r = 0; r < R; r++)
#pragma unroll
            for (int b = 0; b < NB; b++) acc[m][r][b] = 0.f;

    for (int u = threadIdx.x; u < units; u += THREADS) {
        WUnit<TYPE> w[NM][R];
#pragma unroll
        for (int m = 0; m < NM; m++)
#pragma unroll
            for (int r = 0; r < R; r++) {
                int row = min(row0 + r, sg.rows - 1);  // clamp: tail rows recompute the last row, never stored
                load_w<TYPE>(m ? sg.qs2 : sg.qs, m ? sg.d2 : sg.d, (int64_t)row * units, (int64_t)row * blocks_per_row, u, w[m][r]);
            }
#pragma unroll
        for (int b = 0; b < NB; b++) {
            float xr[UE];
            load_x<TYPE>(a.x + (int64_t)b * a.x_stride, u, xr);
#pragma unroll
            for (int m = 0; m < NM; m++)
#pragma unroll
                for (int r = 0; r < R; r++) acc[m][r][b] += dot_unit<TYPE>(w[m][r], xr);
        }
    }

    // warp reduce, then across warps through smem
    constexpr int NW = THREADS / 32;
    __shared__ float red[NW][NM * R * NB];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int m = 0; m < NM; m++)
#pragma unroll
        for (int r = 0; r < R; r++)
#pragma unroll
            for (int b = 0; b < NB; b++) {
                float v = warp_sum(acc[m][r][b]);
                if (lane == 0) red[warp][(m * R + r) * NB + b] = v;
            }
    __syncthreads();
    if (threadIdx.x < R * NB) {
        const int r = threadIdx.x / NB, b = threadIdx.x % NB, row = row0 + r;
        if (row < sg.rows) {
            float v = 0.f, v2 = 0.f;
#pragma unroll
            for (int wi = 0; wi < NW; wi++) {
                v += red[wi][r * NB + b];
                if (NM == 2) v2 += red[wi][(R + r) * NB + b];
            }
            if (sg.bias) v += sg.bias[row];
            float *o = sg.out + (int64_t)b * sg.out_stride + row;
            if (EPI == EPI_STORE) *o = v;
            else if (EPI == EPI_RESID) *o += v;           // X += W·x, go/model.go:592-594, :610-612
            else *o = silu_f(v) * v2;                      // SiLU(gate) * up, go/model.go:604-606
        }
    }
}

}  // namespace nl
