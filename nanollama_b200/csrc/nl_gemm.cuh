// nl_gemm.cuh — prefill GEMM on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM), sm_100a.
//
//   C[T, N] (=|+=) A[T, K] · W[N, K]^T        the T-token form of matmulDispatch (go/model.go:361-386); the reference feeds a
//                                             prompt one Forward at a time (go/main.go:160-166), this does it in one pass.
//
// W stays in the library's planar GGUF layout in HBM and is dequantised on the fly into shared memory as bf16, A arrives as
// bf16.  To keep the logits inside the 1e-3 gate both operands are split into two bf16 terms (w = w_hi + w_lo is EXACT:
// (n-8)*d has <= 15 significant bits; a = a_hi + a_lo keeps 16) and three products are accumulated in fp32 in TMEM:
// a_hi*w_hi + a_hi*w_lo + a_lo*w_hi  (the dropped a_lo*w_lo term is ~2^-17 relative).
//
// One CTA computes a 128 x 128 tile of C.  8 warps fill a 2-stage shared-memory ring (K step 64) in the canonical K-major
// no-swizzle UMMA layout (8x8 core matrices, 128 B each); one thread issues tcgen05.mma (M=128, N=128, K=16) and commits each
// stage to an mbarrier; the epilogue reads the accumulator with tcgen05.ld (32 lanes x 32 columns per warp).
#pragma once
#include <cuda_bf16.h>

#include "nl_common.cuh"
#include "nl_stream.cuh"  // mbarrier wrappers

namespace nl {

constexpr int GM_BM = 128, GM_BN = 128, GM_BK = 64;
constexpr int GM_THREADS = 256;
constexpr int GM_TILE_BYTES = GM_BM * GM_BK * 2;             // one bf16 operand tile: 16 KB
constexpr int GM_MAX_MT = 2;                                 // M tiles (128 rows each) per CTA sharing one dequantised W tile
constexpr int GM_STAGES = 2;
__host__ __device__ constexpr int gm_stage_bytes(int mt) { return (2 * mt + 2) * GM_TILE_BYTES; }   // mt x (a_hi, a_lo), w_hi, w_lo
constexpr int GM_LBO = (GM_BM / 8) * 128;                    // bytes between core matrices adjacent in K   (2048)
constexpr int GM_SBO = 128;                                  // bytes between core matrices adjacent in M/N

enum { GEPI_STORE = 0, GEPI_RESID = 1 };

struct GemmArgs {
    const __nv_bfloat16 *a_hi, *a_lo;  // [T][K] row-major
    const uint8_t *qs;                 // planar quants (or raw F16 rows)
    const __half *d;                   // block scales
    const float *bias;                 // [N] or null
    float *c;                          // [T][ldc]
    int T, N, K, ldc, epi;
    int swap_lbo_sbo;                  // debug switch for the descriptor convention
};

__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
    // sm_100 shared-memory matrix descriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), no swizzle
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
// instruction descriptor kind::f16: D=f32 [4,6)=1, A=bf16 [7,10)=1, B=bf16 [10,13)=1, both K-major, N>>3 [17,23), M>>4 [24,29)
constexpr uint32_t GM_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((GM_BN >> 3) << 17) | ((GM_BM >> 4) << 24);

__device__ __forceinline__ void umma_f16(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_c), "l"(adesc),
                 "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// split an fp32 pair into packed bf16 hi and the exact-remainder lo
__device__ __forceinline__ void split2(float x, float y, uint32_t &hi, uint32_t &lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(x, y);          // one packed conversion
    hi = *reinterpret_cast<const uint32_t *>(&h);
    const float fx = __uint_as_float(hi << 16), fy = __uint_as_float(hi & 0xFFFF0000u);   // bf16 -> fp32 is a shift
    const __nv_bfloat162 l = __floats2bfloat162_rn(x - fx, y - fy);
    lo = *reinterpret_cast<const uint32_t *>(&l);
}

// byte offset of the 16-byte chunk (row r, k-chunk kc) inside an operand tile
__device__ __forceinline__ uint32_t tile_off(int r, int kc) { return (uint32_t)kc * GM_LBO + (uint32_t)(r >> 3) * GM_SBO + (uint32_t)(r & 7) * 16; }

template <int TYPE, int MT>
__global__ void __launch_bounds__(GM_THREADS, 1) gemm_tc_kernel(const GemmArgs g) {
    constexpr int STAGE_BYTES = gm_stage_bytes(MT);
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t mma_bar[GM_STAGES];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n0 = blockIdx.x * GM_BN, m0 = blockIdx.y * GM_BM * MT;
    const int nb = g.K / 32;

    if (tid == 0) {
        for (int s = 0; s < GM_STAGES; s++) mbar_init(&mma_bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {  // 128 TMEM columns per fp32 128x128 accumulator
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(128 * MT) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_c = tmem_base_s;

    const int ksteps = g.K / GM_BK;
    // raw operands of one K step, fetched one step ahead so that the global-load latency hides behind the conversion of the
    // previous step and the MMAs in flight
    struct Raw { uint4 wq[4]; float d; };
    const int wn = tid & 127, wb = tid >> 7;            // W: my row inside the tile and my quant block inside the K step
    const bool w_ok = n0 + wn < g.N;
    auto fetch = [&](int ks, Raw &r) {
        const int k0 = ks * GM_BK;
        r.d = 0.f;
#pragma unroll
        for (int i = 0; i < 4; i++) r.wq[i] = make_uint4(0, 0, 0, 0);
        if (w_ok) {
            if constexpr (TYPE == NL_Q4_0) {
                const size_t bi = (size_t)(n0 + wn) * nb + (k0 >> 5) + wb;
                r.wq[0] = ldg_stream_u4(g.qs + bi * 16);
                r.d = __half2float(g.d[bi]);
            } else if constexpr (TYPE == NL_Q8_0) {
                const size_t bi = (size_t)(n0 + wn) * nb + (k0 >> 5) + wb;
                r.wq[0] = ldg_stream_u4(g.qs + bi * 32);
                r.wq[1] = ldg_stream_u4(g.qs + bi * 32 + 16);
                r.d = __half2float(g.d[bi]);
            } else {
                const __half *wp = reinterpret_cast<const __half *>(g.qs) + (size_t)(n0 + wn) * g.K + k0 + 32 * wb;
#pragma unroll
                for (int i = 0; i < 4; i++) r.wq[i] = ldg_stream_u4(wp + 8 * i);
            }
        }
    };
    Raw cur;
    fetch(0, cur);
    for (int ks = 0; ks < ksteps; ks++) {
        const int s = ks & 1;
        Raw nxt;
        if (ks + 1 < ksteps) fetch(ks + 1, nxt);
        if (ks >= GM_STAGES) mbar_wait(&mma_bar[s], ((ks >> 1) - 1) & 1);  // the MMAs that read this stage have completed
        uint8_t *st = smem + (size_t)s * STAGE_BYTES;
        uint8_t *w_hi = st + 2 * MT * GM_TILE_BYTES, *w_lo = w_hi + GM_TILE_BYTES;
        // ---- A: MT x 128 rows x 8 chunks of 8 bf16, two planes each: cp.async straight into the UMMA layout (rows >= T: zeros);
        //      the copies land while this thread dequantises its share of W below
        {
            const int k0 = ks * GM_BK;
#pragma unroll
            for (int j = 0; j < 4 * MT; j++) {
                const int q = tid + GM_THREADS * j, row = q & (128 * MT - 1), kc = q / (128 * MT), mt = row >> 7, r = row & 127;
                uint8_t *dh = st + (2 * mt) * GM_TILE_BYTES + tile_off(r, kc), *dl = dh + GM_TILE_BYTES;
                if (m0 + row < g.T) {
                    const size_t off = (size_t)(m0 + row) * g.K + k0 + kc * 8;
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dh)), "l"(g.a_hi + off) : "memory");
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dl)), "l"(g.a_lo + off) : "memory");
                } else {
                    *reinterpret_cast<uint4 *>(dh) = make_uint4(0, 0, 0, 0);
                    *reinterpret_cast<uint4 *>(dl) = make_uint4(0, 0, 0, 0);
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        // ---- W: thread = (row n, quant block b of this K step); dequantise 32 weights, split, store 4 chunks per plane
        {
            float w[32];
            if constexpr (TYPE == NL_Q4_0) {
                const uint32_t ws[4] = {cur.wq[0].x, cur.wq[0].y, cur.wq[0].z, cur.wq[0].w};
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const uint32_t lo4 = ws[i] & 0x0F0F0F0Fu, hi4 = (ws[i] >> 4) & 0x0F0F0F0Fu;
#pragma unroll
                    for (int k = 0; k < 4; k++) {   // exact int -> fp32 without I2F: 0x4B000000 | n is 2^23 + n
                        w[4 * i + k] = (u8_to_f32_magic(lo4, k) - 8388616.0f) * cur.d;
                        w[4 * i + k + 16] = (u8_to_f32_magic(hi4, k) - 8388616.0f) * cur.d;
                    }
                }
            } else if constexpr (TYPE == NL_Q8_0) {
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const uint32_t ws[4] = {cur.wq[h].x ^ 0x80808080u, cur.wq[h].y ^ 0x80808080u, cur.wq[h].z ^ 0x80808080u, cur.wq[h].w ^ 0x80808080u};
#pragma unroll
                    for (int i = 0; i < 4; i++)
#pragma unroll
                        for (int k = 0; k < 4; k++) w[16 * h + 4 * i + k] = (u8_to_f32_magic(ws[i], k) - 8388736.0f) * cur.d;
                }
            } else {
#pragma unroll
                for (int h = 0; h < 4; h++) {
                    const uint32_t ws[4] = {cur.wq[h].x, cur.wq[h].y, cur.wq[h].z, cur.wq[h].w};
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&ws[i]));
                        w[8 * h + 2 * i] = f.x; w[8 * h + 2 * i + 1] = f.y;
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < 4; c++) {
                uint4 vh, vl;
                split2(w[8 * c + 0], w[8 * c + 1], vh.x, vl.x); split2(w[8 * c + 2], w[8 * c + 3], vh.y, vl.y);
                split2(w[8 * c + 4], w[8 * c + 5], vh.z, vl.z); split2(w[8 * c + 6], w[8 * c + 7], vh.w, vl.w);
                const uint32_t off = tile_off(wn, wb * 4 + c);
                *reinterpret_cast<uint4 *>(w_hi + off) = vh;
                *reinterpret_cast<uint4 *>(w_lo + off) = vl;
            }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the tensor core (async proxy)
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t lbo = g.swap_lbo_sbo ? GM_SBO : GM_LBO, sbo = g.swap_lbo_sbo ? GM_LBO : GM_SBO;
#pragma unroll
            for (int kk = 0; kk < GM_BK / 16; kk++) {   // one MMA consumes K=16 = two core matrices along K
                const uint32_t koff = kk * 2 * GM_LBO;
                const uint64_t wh = umma_desc(smem_u32(w_hi) + koff, lbo, sbo), wl = umma_desc(smem_u32(w_lo) + koff, lbo, sbo);
#pragma unroll
                for (int mt = 0; mt < MT; mt++) {
                    const uint32_t a_hi = smem_u32(st + (2 * mt) * GM_TILE_BYTES), a_lo = a_hi + GM_TILE_BYTES;
                    const uint64_t ah = umma_desc(a_hi + koff, lbo, sbo), al = umma_desc(a_lo + koff, lbo, sbo);
                    const uint32_t acc = tmem_c + mt * 128;   // accumulator mt lives in TMEM columns [128 mt, 128 mt + 128)
                    umma_f16(acc, ah, wh, GM_IDESC, (ks | kk) != 0);
                    umma_f16(acc, ah, wl, GM_IDESC, 1);
                    umma_f16(acc, al, wh, GM_IDESC, 1);
                }
            }
            umma_commit(&mma_bar[s]);   // arrives when every MMA issued so far has finished (implies fence::before_thread_sync)
        }
        cur = nxt;
    }
    // ---- epilogue: wait for the last commit (it covers all earlier MMAs), TMEM -> registers -> global
    {
        const int last = ksteps - 1;
        mbar_wait(&mma_bar[last & 1], (last >> 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int lg = warp & 3, ch = warp >> 2;           // TMEM lane group of this warp, column half
#pragma unroll
        for (int mc = 0; mc < 2 * MT; mc++) {
            const int mt = mc >> 1, cc = mc & 1;
            const int row = m0 + mt * 128 + lg * 32 + lane;
            const int col0 = ch * 64 + cc * 32;
            uint32_t r[32];
            const uint32_t taddr = tmem_c + ((uint32_t)(lg * 32) << 16) + (uint32_t)(mt * 128 + col0);
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,"
                "%28,%29,%30,%31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
                  "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
                  "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (row < g.T) {
                float *cp = g.c + (size_t)row * g.ldc + n0 + col0;
#pragma unroll
                for (int i = 0; i < 32; i++) {
                    const int n = n0 + col0 + i;
                    if (n < g.N) {
                        float v = __uint_as_float(r[i]);
                        if (g.bias) v += g.bias[n];
                        if (g.epi == GEPI_RESID) v += cp[i];
                        cp[i] = v;
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_c), "n"(128 * MT) : "memory");
}

// fp32 [n] -> bf16 hi / lo planes (exact two-term split up to 16 bits)
static __global__ void split_bf16_kernel(const float *__restrict__ x, __nv_bfloat16 *__restrict__ hi, __nv_bfloat16 *__restrict__ lo, int64_t n, int K, int tiled) {
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
    if (i + 1 < n) {
        uint32_t h, l;
        split2(x[i], x[i + 1], h, l);
        const size_t o = tiled ? plane_index((int)(i / K), (int)(i % K), K, 1) : (size_t)i;   // (tiled: K is a multiple of 32)
        *reinterpret_cast<uint32_t *>(hi + o) = h;
        *reinterpret_cast<uint32_t *>(lo + o) = l;
    } else if (i < n) {
        const __nv_bfloat16 hx = __float2bfloat16_rn(x[i]);
        hi[i] = hx; lo[i] = __float2bfloat16_rn(x[i] - __bfloat162float(hx));
    }
}

template <int TYPE, int MT>
static int launch_gemm_mt(const GemmArgs &g, cudaStream_t st) {
    static bool configured = false;
    auto kern = gemm_tc_kernel<TYPE, MT>;
    const size_t smem = (size_t)GM_STAGES * gm_stage_bytes(MT);
    if (!configured) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -2;
        configured = true;
    }
    dim3 grid((g.N + GM_BN - 1) / GM_BN, (g.T + GM_BM * MT - 1) / (GM_BM * MT));
    kern<<<grid, GM_THREADS, smem, st>>>(g);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
template <int TYPE>
int launch_gemm_typed(const GemmArgs &g, cudaStream_t st) {
    // two 128-token tiles per CTA share one dequantised W tile once there are enough tokens to fill them
    return g.T > GM_BM ? launch_gemm_mt<TYPE, 2>(g, st) : launch_gemm_mt<TYPE, 1>(g, st);
}
int launch_gemm_q4_0(const GemmArgs &g, cudaStream_t st);
int launch_gemm_q8_0(const GemmArgs &g, cudaStream_t st);
int launch_gemm_f16(const GemmArgs &g, cudaStream_t st);

}  // namespace nl
