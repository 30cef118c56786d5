// nl_common.cuh — shared types/helpers for libnanollama_cuda.so (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/nanollama_cuda.h"

namespace nl {

// ---- error plumbing (thread-local message behind nl_last_error) ----
void set_error(const char *fmt, ...);
int fail(int code, const char *fmt, ...);

#define NL_CUDA(expr)                                                                                   \
    do {                                                                                                \
        cudaError_t _e = (expr);                                                                        \
        if (_e != cudaSuccess) return nl::fail(NL_ERR_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

// ---- GGML block geometry (go/gguf.go:238-272) ----
__host__ __device__ inline int blk_elems(int t) { return (t == NL_F32 || t == NL_F16) ? 1 : (t == NL_Q4_K || t == NL_Q6_K) ? 256 : 32; }
__host__ __device__ inline int blk_bytes(int t) {
    switch (t) {
    case NL_F32: return 4; case NL_F16: return 2; case NL_Q4_0: return 18; case NL_Q5_0: return 22;
    case NL_Q8_0: return 34; case NL_Q4_K: return 144; case NL_Q6_K: return 210; default: return 0;
    }
}
inline int64_t tensor_nbytes(int t, int64_t n) { return blk_bytes(t) ? n / blk_elems(t) * blk_bytes(t) : -1; }

// ---- device-resident weight matrix in the library's planar layout ----
// The 18-B / 34-B GGUF blocks are only 2-B aligned, so at upload each matrix is split (bijectively) into
//   qs : [rows][cols/32] x 16 B (Q4_0 nibbles, byte j = elem j | elem j+16 << 4)  or  32 B (Q8_0 int8)
//   d  : [rows][cols/32] fp16 block scales
// Same bytes as the GGUF tensor, now 16-B aligned for 128-bit loads / bulk copies.  F16/F32 are kept as they are.
// Q5_0/Q4_K/Q6_K are kept in raw GGUF block order (qs = raw bytes).
struct DevMat {
    int type = -1;
    int64_t rows = 0, cols = 0;
    uint8_t *qs = nullptr;
    __half *d = nullptr;
    size_t qs_bytes = 0, d_bytes = 0;
    bool present() const { return qs != nullptr; }
    size_t bytes() const { return qs_bytes + d_bytes; }
};

// Element index of activation (row, k) inside one bf16 plane of a GEMM input: row-major [T][K], or (tiled != 0) the order in which
// nl_gemm2.cuh wants it in shared memory, so that a 128-row x 32-k tile is ONE contiguous 8 KB piece a bulk copy can fetch:
// tiles [row / 128][k / 32], inside a tile the K-major UMMA core-matrix order [k / 8 % 4][row % 128 / 8][row % 8][k % 8].
__host__ __device__ __forceinline__ size_t plane_index(int row, int k, int K, int tiled) {
    if (!tiled) return (size_t)row * K + k;
    return ((size_t)(row >> 7) * (size_t)(K >> 5) + (size_t)(k >> 5)) * 4096 + (size_t)(((k >> 3) & 3) * 1024 + ((row & 127) >> 3) * 64 + (row & 7) * 8 + (k & 7));
}

// ---- small device helpers ----
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// streaming 128-/64-bit loads that do not pollute L1 (weights are read once per token)
__device__ __forceinline__ uint4 ldg_stream_u4(const void *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint2 ldg_stream_u2(const void *p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ unsigned short ldg_stream_u16(const void *p) {
    unsigned short r;
    asm volatile("ld.global.nc.L1::no_allocate.u16 %0, [%1];" : "=h"(r) : "l"(p));
    return r;
}
// exact small-int -> fp32 without I2F: 0x4B000000|n is 2^23+n
__device__ __forceinline__ float u8_to_f32_magic(uint32_t word, int byte_idx) {
    return __uint_as_float(__byte_perm(word, 0x4B000000u, 0x7540u | byte_idx));
}

// binary16 -> binary32 with the exact semantics of the reference's 65,536-entry table (go/gguf.go:601-636): NaNs keep their
// payload (mant << 13) instead of being canonicalised the way cvt.f32.f16 does.  Used where bit-exactness is the contract
// (dequant, embedding rows); the GEMV inner loops use the plain conversion (identical for every non-NaN input).
__device__ __forceinline__ float h2f_exact(unsigned short h) {
    if ((h & 0x7C00u) == 0x7C00u) return __uint_as_float(((uint32_t)(h & 0x8000u) << 16) | 0x7F800000u | ((uint32_t)(h & 0x03FFu) << 13));
    return __half2float(__ushort_as_half(h));
}
__device__ __forceinline__ float h2f_exact(__half h) { return h2f_exact(__half_as_ushort(h)); }

__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + expf(-x)); }  // go/quant.go:629-631

}  // namespace nl
