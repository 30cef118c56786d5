// nl_stream.cuh — persistent, bulk-copy-staged, dequant-fused GEMV for batch-1 decode (sm_100a).
//
//   out[row] (=|+=) sum_k W[row][k] * act(x)[k]          matmulDispatch, go/model.go:361-386 (MatMulQ4_0/Q8_0/F16)
//
// One CTA per SM slot streams a contiguous band of rows.  A producer warp moves the band through a ring of shared-memory
// stages with cp.async.bulk (1-D TMA bulk copies) completing on mbarriers, so a fixed number of bytes is always in flight
// per SM no matter what the math warps are doing.  12 consumer warps own one 32-element block COLUMN each and keep that
// slice of x in registers for the whole kernel (x is constant across rows), so shared memory only carries weights.
// Per-row partials are folded with a packed butterfly, dropped into a per-stage scratch, and the otherwise idle producer
// warp finishes them (fixed summation order => deterministic), applying bias / residual / SwiGLU.
// Fusions: RMSNorm of the input (RMSNormInto, go/quant.go:597-607) in the prologue; residual add (go/model.go:592-594);
// SiLU(gate)*up (go/model.go:604-606) by streaming the gate and up rows of the same index together.
// Programmatic dependent launch: weights do not depend on the previous kernel, so the ring is filled BEFORE
// griddepcontrol.wait and the previous kernel's tail / the next kernel's prologue overlap.
#pragma once
#include "nl_common.cuh"

namespace nl {

constexpr int ST_CONSUMER_WARPS = 12;
constexpr int ST_CONSUMERS = ST_CONSUMER_WARPS * 32;  // 384
constexpr int ST_THREADS = ST_CONSUMERS + 32;         // + producer warp
constexpr int ST_MAX_STAGES = 8;
#ifndef NL_ST_MIN_CTAS
#define NL_ST_MIN_CTAS 1
#endif
constexpr int ST_MIN_CTAS = NL_ST_MIN_CTAS;            // 2: a PDL successor's CTAs can co-reside and prefetch weights during this kernel's tail
constexpr int ST_RED = 48;                            // max (rows per tile) * (warps per row)

enum { ACT_NONE = 0, ACT_RMSNORM = 1 };
enum { SEPI_STORE = 0, SEPI_RESID = 1, SEPI_SWIGLU = 2 };

struct StreamSeg {
    const uint8_t *qs, *qs2;  // planar quants (qs2: the "up" matrix for SEPI_SWIGLU)
    const __half *d, *d2;     // block scales
    const float *bias;
    float *out;
    int rows;
    int tile_begin;           // first global tile index of this segment
};
struct StreamArgs {
    StreamSeg seg[3];
    int nseg;
    int total_tiles;
    const float *x;           // [cols] input activations
    const float *norm_w;      // ACT_RMSNORM: weight [cols]
    float eps;
    int cols, nb, nb_pad;     // nb = cols/32, nb_pad = nb rounded up to 32
    int RG;                   // row groups = ST_CONSUMERS / nb_pad
    int T;                    // rows per tile = RPT * RG
    int stages;
    int stage_bytes;          // bytes of one ring stage (all matrices of the tile)
    int epi;                  // SEPI_*
};

// ---- PTX wrappers ----
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)), "l"(src_gmem),
                 "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s_hint(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, %0;" ::"n"(ST_CONSUMERS) : "memory"); }

template <int TYPE> struct BlkBytes;
template <> struct BlkBytes<NL_Q4_0> { static constexpr int Q = 16, D = 2; };
template <> struct BlkBytes<NL_Q8_0> { static constexpr int Q = 32, D = 2; };
template <> struct BlkBytes<NL_F16> { static constexpr int Q = 64, D = 0; };

// dot of one 32-element block with the register-resident x slice, 4 independent accumulators
// Q4_0: nibble n -> fp32 (32 + 2n) with a shift+mask per 4 nibbles and one PRMT each (mantissa bits 19..22 of 32.0f); the
//       caller adds -48*sum(x) per block (offset 32 + zero-point 2*8) and halves the scale.  Exact products, fp32 accumulate.
template <int TYPE>
__device__ __forceinline__ float block_dot(const uint8_t *qp, const float (&xr)[32]) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    if constexpr (TYPE == NL_Q4_0) {
        const uint4 q = *reinterpret_cast<const uint4 *>(qp);
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int i = 0; i < 4; i++) {
            // byte = n<<3 lands in mantissa bits 19..22 of 0x42000000 (=32.0): value 32 + 2n, no OR needed
            const uint32_t lo = (w[i] << 3) & 0x78787878u;
            const uint32_t hi = (w[i] >> 1) & 0x78787878u;
            a0 = fmaf(__uint_as_float(__byte_perm(lo, 0x42000000u, 0x7044u)), xr[4 * i + 0], a0);
            a1 = fmaf(__uint_as_float(__byte_perm(lo, 0x42000000u, 0x7144u)), xr[4 * i + 1], a1);
            a2 = fmaf(__uint_as_float(__byte_perm(lo, 0x42000000u, 0x7244u)), xr[4 * i + 2], a2);
            a3 = fmaf(__uint_as_float(__byte_perm(lo, 0x42000000u, 0x7344u)), xr[4 * i + 3], a3);
            a0 = fmaf(__uint_as_float(__byte_perm(hi, 0x42000000u, 0x7044u)), xr[16 + 4 * i + 0], a0);
            a1 = fmaf(__uint_as_float(__byte_perm(hi, 0x42000000u, 0x7144u)), xr[16 + 4 * i + 1], a1);
            a2 = fmaf(__uint_as_float(__byte_perm(hi, 0x42000000u, 0x7244u)), xr[16 + 4 * i + 2], a2);
            a3 = fmaf(__uint_as_float(__byte_perm(hi, 0x42000000u, 0x7344u)), xr[16 + 4 * i + 3], a3);
        }
    } else if constexpr (TYPE == NL_Q8_0) {
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const uint4 q = *reinterpret_cast<const uint4 *>(qp + 16 * h);
            const uint32_t w[4] = {q.x ^ 0x80808080u, q.y ^ 0x80808080u, q.z ^ 0x80808080u, q.w ^ 0x80808080u};
#pragma unroll
            for (int i = 0; i < 4; i++) {
                a0 = fmaf(u8_to_f32_magic(w[i], 0) - 8388736.0f, xr[16 * h + 4 * i + 0], a0);
                a1 = fmaf(u8_to_f32_magic(w[i], 1) - 8388736.0f, xr[16 * h + 4 * i + 1], a1);
                a2 = fmaf(u8_to_f32_magic(w[i], 2) - 8388736.0f, xr[16 * h + 4 * i + 2], a2);
                a3 = fmaf(u8_to_f32_magic(w[i], 3) - 8388736.0f, xr[16 * h + 4 * i + 3], a3);
            }
        }
    } else {
#pragma unroll
        for (int h = 0; h < 4; h++) {
            const uint4 q = *reinterpret_cast<const uint4 *>(qp + 16 * h);
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&w[i]));
                if (i & 1) { a2 = fmaf(f.x, xr[8 * h + 2 * i], a2); a3 = fmaf(f.y, xr[8 * h + 2 * i + 1], a3); }
                else { a0 = fmaf(f.x, xr[8 * h + 2 * i], a0); a1 = fmaf(f.y, xr[8 * h + 2 * i + 1], a1); }
            }
        }
    }
    return (a0 + a1) + (a2 + a3);
}

// packed butterfly: RPT per-lane partials -> lane ((idx) * 32/RPT) holds the warp total of partial idx
template <int RPT>
__device__ __forceinline__ float warp_fold(const float (&p)[RPT], int lane) {
    float k;
    if constexpr (RPT == 4) {
        const bool up = lane & 16;
        float k0 = up ? p[2] : p[0], k1 = up ? p[3] : p[1];
        k0 += __shfl_xor_sync(0xffffffffu, up ? p[0] : p[2], 16);
        k1 += __shfl_xor_sync(0xffffffffu, up ? p[1] : p[3], 16);
        const bool up2 = lane & 8;
        k = up2 ? k1 : k0;
        k += __shfl_xor_sync(0xffffffffu, up2 ? k0 : k1, 8);
    } else if constexpr (RPT == 2) {
        const bool up = lane & 16;
        k = up ? p[1] : p[0];
        k += __shfl_xor_sync(0xffffffffu, up ? p[0] : p[1], 16);
        k += __shfl_xor_sync(0xffffffffu, k, 8);
    } else {
        k = p[0];
        k += __shfl_xor_sync(0xffffffffu, k, 16);
        k += __shfl_xor_sync(0xffffffffu, k, 8);
    }
    k += __shfl_xor_sync(0xffffffffu, k, 4);
    k += __shfl_xor_sync(0xffffffffu, k, 2);
    k += __shfl_xor_sync(0xffffffffu, k, 1);
    return k;  // RPT=4: lanes 0,8,16,24 -> partial 0,1,2,3;  RPT=2: lanes 0,16 -> 0,1;  RPT=1: lane 0
}

template <int TYPE, int NM, int RPT, int ACT>
__global__ void __launch_bounds__(ST_THREADS, ST_MIN_CTAS) gemv_stream_kernel(const StreamArgs a) {
    constexpr int QB = BlkBytes<TYPE>::Q, DB = BlkBytes<TYPE>::D;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t full_bar[ST_MAX_STAGES], empty_bar[ST_MAX_STAGES];
    __shared__ float red[ST_MAX_STAGES][NM][ST_RED];
    __shared__ double ss_red[ST_CONSUMER_WARPS];
    __shared__ float inv_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int NS = a.stages, T = a.T, nb = a.nb;
    // contiguous band of tiles for this CTA
    const int t_begin = (int)(((long long)a.total_tiles * blockIdx.x) / gridDim.x);
    const int t_end = (int)(((long long)a.total_tiles * (blockIdx.x + 1)) / gridDim.x);
    const int n_my = t_end - t_begin;

    if (tid == 0) {
        for (int s = 0; s < NS; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], ST_CONSUMER_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    pdl_launch_dependents();

    const int q_tile_bytes = T * nb * QB, d_tile_bytes = T * nb * DB;

    auto locate = [&](int g, int &s, int &row0) {
        s = 0;
        if (a.nseg > 1 && g >= a.seg[1].tile_begin) s = 1;
        if (a.nseg > 2 && g >= a.seg[2].tile_begin) s = 2;
        row0 = (g - a.seg[s].tile_begin) * T;
    };

    if (warp == ST_CONSUMER_WARPS) {
        // ===================== producer + finisher warp =====================
        bool waited = false;
        for (int it = 0; it < n_my + NS; it++) {
            const int slot = it % NS, round = it / NS;
            if (it >= NS) {
                mbar_wait(&empty_bar[slot], (round - 1) & 1);  // tile it-NS fully consumed, its partials are in red[slot]
                if (!waited) { pdl_wait(); waited = true; }
                int s, row0;
                locate(t_begin + it - NS, s, row0);
                const StreamSeg &sg = a.seg[s];
                const int wpr = a.nb_pad >> 5;
                for (int r = lane; r < T; r += 32) {
                    const int row = row0 + r;
                    if (row < sg.rows) {
                        float v = 0.f, v2 = 0.f;
                        for (int wi = 0; wi < wpr; wi++) {
                            v += red[slot][0][r * wpr + wi];
                            if (NM == 2) v2 += red[slot][NM - 1][r * wpr + wi];
                        }
                        if (sg.bias) v += sg.bias[row];
                        if (NM == 2) sg.out[row] = silu_f(v) * v2;
                        else if (a.epi == SEPI_RESID) sg.out[row] += v;  // X += W·x, go/model.go:592-594, :610-612
                        else sg.out[row] = v;
                    }
                }
                __syncwarp();
            }
            if (it < n_my && lane == 0) {
                int s, row0;
                locate(t_begin + it, s, row0);
                const StreamSeg &sg = a.seg[s];
                const int rows_here = min(T, sg.rows - row0);
                const uint32_t qb = (uint32_t)rows_here * nb * QB, db = (uint32_t)rows_here * nb * DB;
                uint8_t *st = smem + (size_t)slot * a.stage_bytes;
                mbar_expect_tx(&full_bar[slot], (qb + db) * NM);
                bulk_g2s(st, sg.qs + (size_t)row0 * nb * QB, qb, &full_bar[slot]);
                if (DB) bulk_g2s(st + q_tile_bytes, reinterpret_cast<const uint8_t *>(sg.d) + (size_t)row0 * nb * DB, db, &full_bar[slot]);
                if (NM == 2) {
                    uint8_t *st2 = st + q_tile_bytes + d_tile_bytes;
                    bulk_g2s(st2, sg.qs2 + (size_t)row0 * nb * QB, qb, &full_bar[slot]);
                    if (DB) bulk_g2s(st2 + q_tile_bytes, reinterpret_cast<const uint8_t *>(sg.d2) + (size_t)row0 * nb * DB, db, &full_bar[slot]);
                }
            }
        }
        return;
    }

    // ===================== consumer warps =====================
    const int wpr = a.nb_pad >> 5;        // warps per row group
    const int rg = warp / wpr, wi = warp % wpr;
    const int c = wi * 32 + lane;         // my block column
    const bool active = (rg < a.RG) && (c < nb);

    pdl_wait();                           // x is produced by the previous kernel
    float xr[32];
    if (active) {
        const float4 *xp = reinterpret_cast<const float4 *>(a.x + c * 32);
#pragma unroll
        for (int i = 0; i < 8; i++) { float4 v = xp[i]; xr[4 * i] = v.x; xr[4 * i + 1] = v.y; xr[4 * i + 2] = v.z; xr[4 * i + 3] = v.w; }
    } else {
#pragma unroll
        for (int i = 0; i < 32; i++) xr[i] = 0.f;
    }
    if constexpr (ACT == ACT_RMSNORM) {
        // float64 sum of squares over the whole vector (row group 0 covers every column once), go/quant.go:598-603
        double ss = 0.0;
        if (active && rg == 0) {
#pragma unroll
            for (int i = 0; i < 32; i++) ss += (double)xr[i] * (double)xr[i];
        }
        ss = warp_sum_d(ss);
        if (lane == 0) ss_red[warp] = ss;
        consumer_bar();
        if (warp == 0) {
            double v = lane < ST_CONSUMER_WARPS ? ss_red[lane] : 0.0;
            v = warp_sum_d(v);
            if (lane == 0) inv_s = (float)(1.0 / sqrt(v / (double)a.cols + (double)a.eps));
        }
        consumer_bar();
        const float inv = inv_s;
        if (active) {
            const float4 *wp = reinterpret_cast<const float4 *>(a.norm_w + c * 32);
#pragma unroll
            for (int i = 0; i < 8; i++) {
                float4 w4 = wp[i];
                xr[4 * i] = xr[4 * i] * inv * w4.x; xr[4 * i + 1] = xr[4 * i + 1] * inv * w4.y;
                xr[4 * i + 2] = xr[4 * i + 2] * inv * w4.z; xr[4 * i + 3] = xr[4 * i + 3] * inv * w4.w;
            }
        }
    }
    float xoff = 0.f;  // Q4_0: -(32 + 2*8) * sum(x) removes the float-encoding offset and the zero point in one go
    if constexpr (TYPE == NL_Q4_0) {
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
        for (int i = 0; i < 8; i++) { s0 += xr[4 * i]; s1 += xr[4 * i + 1]; s2 += xr[4 * i + 2]; s3 += xr[4 * i + 3]; }
        xoff = -48.0f * ((s0 + s1) + (s2 + s3));
    }

    for (int it = 0; it < n_my; it++) {
        const int slot = it % NS;
        mbar_wait(&full_bar[slot], (it / NS) & 1);
        const uint8_t *st = smem + (size_t)slot * a.stage_bytes;
        float part[NM][RPT];
#pragma unroll
        for (int m = 0; m < NM; m++) {
            const uint8_t *qs_s = st + m * (q_tile_bytes + d_tile_bytes);
            const uint8_t *d_s = qs_s + q_tile_bytes;
#pragma unroll
            for (int r = 0; r < RPT; r++) {
                float v = 0.f;
                if (active) {
                    const int row_in_tile = rg + r * a.RG;
                    const int bi = row_in_tile * nb + c;
                    v = block_dot<TYPE>(qs_s + (size_t)bi * QB, xr) + xoff;
                    if (DB) v *= (TYPE == NL_Q4_0 ? 0.5f : 1.0f) * __half2float(*reinterpret_cast<const __half *>(d_s + (size_t)bi * 2));
                }
                part[m][r] = v;
            }
        }
#pragma unroll
        for (int m = 0; m < NM; m++) {
            const float k = warp_fold<RPT>(part[m], lane);
            if ((lane & (32 / RPT - 1)) == 0 && rg < a.RG) {
                const int r = lane / (32 / RPT);
                red[slot][m][(rg + r * a.RG) * wpr + wi] = k;
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[slot]);
    }
}


// ---- host side: plan + launch -------------------------------------------------------------------
struct StreamPlan { bool ok; int nb, nb_pad, RG, RPT, T, stages, stage_bytes, grid; size_t smem; };

// Decide whether the streaming kernel can take this GEMV and how it is tiled.  rows[] are the segment row counts.
inline StreamPlan plan_stream(int type, const int *rows, int nseg, int cols, int NM, int num_sms) {
    StreamPlan p{};
    p.ok = false;
    if (type != NL_Q4_0 && type != NL_Q8_0 && type != NL_F16) return p;
    if (cols % 32) return p;
    const int QB = type == NL_Q4_0 ? 16 : type == NL_Q8_0 ? 32 : 64, DB = type == NL_F16 ? 0 : 2;
    p.nb = cols / 32;
    p.nb_pad = (p.nb + 31) / 32 * 32;
    if (p.nb_pad > ST_CONSUMERS) return p;
    p.RG = ST_CONSUMERS / p.nb_pad;
    const int group_bytes = p.RG * p.nb * (QB + DB) * NM;
    p.RPT = 4 * group_bytes <= 32 * 1024 ? 4 : 2 * group_bytes <= 32 * 1024 ? 2 : 1;
    if (p.RPT * group_bytes > 56 * 1024) return p;
    p.T = p.RPT * p.RG;
    if (DB && ((p.T * p.nb * DB) % 16)) return p;                       // bulk-copy size/alignment of the scale plane
    for (int i = 0; i < nseg; i++) if (DB && (((rows[i] % p.T) * p.nb * DB) % 16)) return p;
    p.stage_bytes = (NM * p.T * p.nb * (QB + DB) + 127) / 128 * 128;
    p.stages = (108 * 1024) / p.stage_bytes;
    if (p.stages > ST_MAX_STAGES) p.stages = ST_MAX_STAGES;
    if (p.stages < 2) return p;
    int tiles = 0;
    for (int i = 0; i < nseg; i++) tiles += (rows[i] + p.T - 1) / p.T;
    p.grid = tiles < num_sms ? tiles : num_sms;
    p.smem = (size_t)p.stages * p.stage_bytes;
    p.ok = true;
    return p;
}

template <int TYPE, int NM, int RPT, int ACT>
static int launch_stream_inst(const StreamArgs &a, int grid, size_t smem, cudaStream_t st, bool pdl) {
    static bool configured = false;  // per instantiation
    auto kern = gemv_stream_kernel<TYPE, NM, RPT, ACT>;
    if (!configured) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024) != cudaSuccess) return -2;
        configured = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(ST_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, a) == cudaSuccess ? 0 : -2;
}
template <int TYPE, int NM, int RPT>
static int launch_stream_act(const StreamArgs &a, int act, int grid, size_t smem, cudaStream_t st, bool pdl) {
    return act == ACT_RMSNORM ? launch_stream_inst<TYPE, NM, RPT, ACT_RMSNORM>(a, grid, smem, st, pdl)
                              : launch_stream_inst<TYPE, NM, RPT, ACT_NONE>(a, grid, smem, st, pdl);
}
template <int TYPE>
int launch_stream_typed(const StreamArgs &a, int NM, int RPT, int act, int grid, size_t smem, cudaStream_t st, bool pdl) {
    if (NM == 2) {
        switch (RPT) {
        case 4: return launch_stream_act<TYPE, 2, 4>(a, act, grid, smem, st, pdl);
        case 2: return launch_stream_act<TYPE, 2, 2>(a, act, grid, smem, st, pdl);
        default: return launch_stream_act<TYPE, 2, 1>(a, act, grid, smem, st, pdl);
        }
    }
    switch (RPT) {
    case 4: return launch_stream_act<TYPE, 1, 4>(a, act, grid, smem, st, pdl);
    case 2: return launch_stream_act<TYPE, 1, 2>(a, act, grid, smem, st, pdl);
    default: return launch_stream_act<TYPE, 1, 1>(a, act, grid, smem, st, pdl);
    }
}
int launch_stream_q4_0(const StreamArgs &a, int NM, int RPT, int act, int grid, size_t smem, cudaStream_t st, bool pdl);
int launch_stream_q8_0(const StreamArgs &a, int NM, int RPT, int act, int grid, size_t smem, cudaStream_t st, bool pdl);
int launch_stream_f16(const StreamArgs &a, int NM, int RPT, int act, int grid, size_t smem, cudaStream_t st, bool pdl);

}  // namespace nl
