// nl_mega.cuh — one persistent kernel per decoded token (batch 1): the whole Forward of go/model.go:490-620.
//
// 148 CTAs (one per SM, cooperative launch) walk the same list of PHASES:
//     per layer:  QKV gemv (RMSNorm fused) | attention (RoPE, KV write, split-KV softmax·V) | O gemv (+residual)
//                 | gate/up gemv (RMSNorm fused, SiLU·up epilogue) | down gemv (+residual)
//     finally:    LM-head gemv (final RMSNorm fused)
// separated by grid-wide arrive/poll barriers in global memory instead of kernel boundaries.  Inside every CTA a producer warp
// streams that CTA's share of ALL weight matrices, in phase order, through one shared-memory ring with cp.async.bulk + mbarriers.
// Weights do not depend on activations, so the producer never waits for a grid barrier: while the math warps sit at a phase
// boundary the ring keeps filling with the next matrix and HBM stays busy.  Math warps own one 32-element block column each and
// keep that slice of the phase's input vector in registers (x is constant across rows).  The producer warp doubles as finisher:
// it sums the per-warp partials of a consumed stage in a fixed order (deterministic) and applies bias / residual / SiLU·up.
#pragma once
#include "nl_common.cuh"
#include "nl_stream.cuh"  // PTX wrappers, StreamSeg, BlkBytes

namespace nl {

constexpr int MG_CONSUMER_WARPS = 24;
constexpr int MG_CONSUMERS = MG_CONSUMER_WARPS * 32;  // 768
constexpr int MG_THREADS = MG_CONSUMERS + 32;         // + producer/finisher warp
constexpr int MG_MAX_STAGES = 8;
constexpr int MG_RED = 96;                            // 4 partials per thread and chunk x (RG * warps-per-row <= 24)
constexpr int MG_MAX_GROUP = 8;                       // q heads per kv head handled by one attention item
constexpr int MG_MAX_SPLIT = 16;
constexpr int MG_REPS = 8;                            // activation vectors are kept in this many copies so that 148 CTAs do not
                                                      // all queue on the same L2 lines when a phase starts
constexpr bool MG_GATE_PREFETCH = false;               // true: request a phase's weights only after its input loads were issued
                                                      // (measured slower on B200: the ring should run ahead across phase boundaries)
constexpr int MG_XS_FLOATS = 4096;                     // phase inputs up to this length are shared between row groups via smem
constexpr int MG_ATT_CHUNK = 256;                     // max positions per attention item pass buffer

enum { PH_GEMV = 0, PH_ATTN = 1 };
enum { XS_PLAIN = 0, XS_RMSNORM = 1, XS_ATTN = 2 };

// Work of a GEMV phase is counted in UNITS of RG consecutive rows (one row per row group of math warps).  Every CTA owns a
// contiguous range of units (balanced to +-1 unit); it moves them through the ring in CHUNKS of up to `upc` units, so the
// transfer size (tens of KB) is independent of how finely the rows are balanced across the 148 CTAs.
struct MegaPhase {
    int kind;
    // ---- PH_GEMV ----
    StreamSeg seg[3];     // seg[i].tile_begin = first unit of segment i
    int nseg, total_units;
    int seg_units[3];     // units per segment (ceil(rows / RG)), so the device never divides
    int upc_log2;
    const float *x;       // XS_PLAIN / XS_RMSNORM input vector [cols] (replica 0)
    int x_reps, x_stride; // copies of x (CTA b reads copy b % x_reps) and floats between copies
    int out_reps, out_stride;  // copies of the output vector the finisher writes (same for every segment)
    const float *norm_w;  // XS_RMSNORM
    int xsrc, epi, NM;
    int upc;              // units per chunk and matrix (NM * upc <= 4), sized so that a chunk stays <= ~56 KB
    int cols, nb, nb_pad, RG;
    int q_chunk_bytes, d_chunk_bytes;  // smem footprint of a full chunk of one matrix
    // ---- PH_ATTN ----
    int layer;
};

struct MegaAttn {
    const float *q, *k, *v;      // fresh projections [H*hd], [kvd], [kvd]
    float *kcache, *vcache;      // [L][S][kvd]
    const float *cos_t, *sin_t;  // [S][hd/2]
    const int32_t *pos;          // device scalar
    float *part_acc;             // [H][nsplit][hd]   un-normalised partial outputs
    float *part_ml;              // [H][nsplit][2]    (running max, sum of exp)
    float *out;                  // [reps][H*hd] combined attention output (written by the last split of each kv head)
    int out_reps, out_stride;
    unsigned int *split_cnt;     // [L][n_kv_heads] arrival counters of the splits, zeroed with the phase barriers
    int n_heads, n_kv_heads, seq_len, qk_norm, conj, nsplit;
    float eps, scale;
};

struct MegaArgs {
    const MegaPhase *phases;
    int n_phases;
    unsigned int *bar;           // [n_phases] arrival counters; zeroed before every launch
    MegaAttn at;
    float eps;
    int stages, slot_bytes;
    int hold_mode;               // 0: ring refill never waits; 1: not between 'phase consumed' and 'next input in registers'; 2: gate per phase
    unsigned long long *trace;   // optional (NL_TRACE): [cta][phase][8] globaltimer stamps for latency forensics
};

// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long pack2(float a, float b) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(__float_as_uint(a)), "r"(__float_as_uint(b)));
    return r;
}
__device__ __forceinline__ unsigned long long pack2u(uint32_t a, uint32_t b) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(a), "r"(b));
    return r;
}
__device__ __forceinline__ void ffma2(unsigned long long &acc, unsigned long long a, unsigned long long b) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}
__device__ __forceinline__ unsigned long long fadd2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ float sum2(unsigned long long v) {
    uint32_t lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
    return __uint_as_float(lo) + __uint_as_float(hi);
}

// dot of one 32-element block with the register-resident x slice (16 packed fp32 pairs), packed FFMA2, 4 chains.
// Q4_0: nibble n -> fp32 (32 + 2n): shift+mask per 4 nibbles, one PRMT each into mantissa bits 19..22 of 32.0f; the caller adds
// -48*sum(x) (offset 32 + zero-point 2*8) and halves the scale.  Products are exact; accumulation is fp32.
template <int TYPE>
__device__ __forceinline__ float block_dot2(const uint8_t *qp, const unsigned long long (&x2)[16]) {
    unsigned long long a0 = 0ull, a1 = 0ull, a2 = 0ull, a3 = 0ull;
    if constexpr (TYPE == NL_Q4_0) {
        const uint4 q = *reinterpret_cast<const uint4 *>(qp);
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint32_t lo = (w[i] << 3) & 0x78787878u;
            const uint32_t hi = (w[i] >> 1) & 0x78787878u;
            ffma2(a0, pack2u(__byte_perm(lo, 0x42000000u, 0x7044u), __byte_perm(lo, 0x42000000u, 0x7144u)), x2[2 * i]);
            ffma2(a1, pack2u(__byte_perm(lo, 0x42000000u, 0x7244u), __byte_perm(lo, 0x42000000u, 0x7344u)), x2[2 * i + 1]);
            ffma2(a0, pack2u(__byte_perm(hi, 0x42000000u, 0x7044u), __byte_perm(hi, 0x42000000u, 0x7144u)), x2[8 + 2 * i]);
            ffma2(a1, pack2u(__byte_perm(hi, 0x42000000u, 0x7244u), __byte_perm(hi, 0x42000000u, 0x7344u)), x2[8 + 2 * i + 1]);
        }
        return sum2(fadd2(a0, a1));
    } else if constexpr (TYPE == NL_Q8_0) {
        const unsigned long long off = pack2(-8388736.0f, -8388736.0f);  // -(2^23 + 128)
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const uint4 q = *reinterpret_cast<const uint4 *>(qp + 16 * h);
            const uint32_t w[4] = {q.x ^ 0x80808080u, q.y ^ 0x80808080u, q.z ^ 0x80808080u, q.w ^ 0x80808080u};
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const unsigned long long f01 = fadd2(pack2u(__byte_perm(w[i], 0x4B000000u, 0x7540u), __byte_perm(w[i], 0x4B000000u, 0x7541u)), off);
                const unsigned long long f23 = fadd2(pack2u(__byte_perm(w[i], 0x4B000000u, 0x7542u), __byte_perm(w[i], 0x4B000000u, 0x7543u)), off);
                if (i & 1) { ffma2(a2, f01, x2[8 * h + 2 * i]); ffma2(a3, f23, x2[8 * h + 2 * i + 1]); }
                else { ffma2(a0, f01, x2[8 * h + 2 * i]); ffma2(a1, f23, x2[8 * h + 2 * i + 1]); }
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
                const unsigned long long f2 = pack2(f.x, f.y);
                if (i == 0) ffma2(a0, f2, x2[4 * h + i]);
                else if (i == 1) ffma2(a1, f2, x2[4 * h + i]);
                else if (i == 2) ffma2(a2, f2, x2[4 * h + i]);
                else ffma2(a3, f2, x2[4 * h + i]);
            }
        }
    }
    return (sum2(a0) + sum2(a1)) + (sum2(a2) + sum2(a3));
}

__device__ __forceinline__ unsigned int ld_acquire(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Grid barrier: every CTA adds 1 to the phase's counter after its last output of that phase (release); one thread per CTA
// polls the counter (acquire).  Counters are zeroed by a memset node in front of the kernel.
__device__ __forceinline__ void phase_arrive(unsigned int *bar, int p) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar + p) : "memory");
}
__device__ __forceinline__ void phase_wait(const unsigned int *bar, int p, unsigned int G) {
    while (ld_acquire(bar + p) < G) { __nanosleep(20); }
}
__device__ __forceinline__ bool mbar_test(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// chunk enumeration shared by the producer and the math warps: next chunk of the CTA's unit range [u, u_end)
struct Chunk { int seg, row0, nunits, rows; };
__device__ __forceinline__ Chunk next_chunk(const MegaPhase &P, int &u, int u_end) {
    Chunk c;
    c.seg = 0;
    if (P.nseg > 1 && u >= P.seg[1].tile_begin) c.seg = 1;
    if (P.nseg > 2 && u >= P.seg[2].tile_begin) c.seg = 2;
    const int seg_end = P.seg[c.seg].tile_begin + P.seg_units[c.seg];
    c.nunits = min(P.upc, min(u_end, seg_end) - u);
    c.row0 = (u - P.seg[c.seg].tile_begin) * P.RG;
    c.rows = min(c.nunits * P.RG, P.seg[c.seg].rows - c.row0);
    u += c.nunits;
    return c;
}
__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define MG_TRACE(p, k) do { if (A.trace) A.trace[((size_t)blockIdx.x * A.n_phases + (p)) * 8 + (k)] = gtime(); } while (0)
template <int NT> __device__ __forceinline__ void cons_bar() { asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory"); }
__device__ __forceinline__ void mega_consumer_bar() { cons_bar<MG_CONSUMERS>(); }

// contiguous band of a phase's tiles for CTA b of G
__device__ __forceinline__ void tile_band(int total, int b, int G, int &t0, int &t1) {
    t0 = (int)(((long long)total * b) / G);
    t1 = (int)(((long long)total * (b + 1)) / G);
}

// ---------------------------------------------------------------------------------------------------------------------
// consumer side of one GEMV phase
template <int TYPE, int NM>
__device__ __forceinline__ void consume_phase(const MegaArgs &A, const MegaPhase &P, uint8_t *smem, uint64_t *full_bar, uint64_t *empty_bar,
                                              float (*red)[MG_RED], double *ss_red, float *inv_s, float *xs, int *hold, int *hold1, int &it, int warp, int lane, int p) {
    constexpr int QB = BlkBytes<TYPE>::Q, DB = BlkBytes<TYPE>::D;
    const int UPC = P.upc;       // units per chunk and matrix; NM * UPC <= 4 partials per thread and chunk
    const int nb = P.nb, wpr = P.nb_pad >> 5;
    const int rg = warp / wpr, wi = warp % wpr;
    const int c = wi * 32 + lane;
    const bool active = (rg < P.RG) && (c < nb);
    int u, u_end;
    tile_band(P.total_units, blockIdx.x, gridDim.x, u, u_end);

    // ---- input slice into registers ----
    float xr[32];
    // Only row group 0 reads the phase input from L2 (all 148 CTAs want the same few KB at the same moment, so every redundant
    // read queues at the same L2 slices); it prepares it and hands it to the other row groups through shared memory.
    const bool share = (P.RG > 1) && (P.cols <= MG_XS_FLOATS);
    const bool loader = active && (rg == 0 || !share);
    if (P.xsrc == XS_ATTN) {
        // combine the split-KV partials of the attention phase: column c covers 32 dims of head c*32/hd
        const MegaAttn &at = A.at;
        constexpr int HD = 64;  // (checked on the host)
        const int h = (c * 32) / HD, d0 = (c * 32) % HD;
#pragma unroll
        for (int i = 0; i < 32; i++) xr[i] = 0.f;
        if (loader) {
            float M = -INFINITY;
            for (int s = 0; s < at.nsplit; s++) M = fmaxf(M, at.part_ml[(h * at.nsplit + s) * 2]);
            float den = 0.f;
            for (int s = 0; s < at.nsplit; s++) {
                const float m = at.part_ml[(h * at.nsplit + s) * 2], l = at.part_ml[(h * at.nsplit + s) * 2 + 1];
                if (l > 0.f) {
                    const float wgt = expf(m - M);
                    den = fmaf(wgt, l, den);
                    const float4 *pa = reinterpret_cast<const float4 *>(at.part_acc + ((size_t)(h * at.nsplit + s)) * HD + d0);
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const float4 v = pa[i];
                        xr[4 * i] = fmaf(wgt, v.x, xr[4 * i]); xr[4 * i + 1] = fmaf(wgt, v.y, xr[4 * i + 1]);
                        xr[4 * i + 2] = fmaf(wgt, v.z, xr[4 * i + 2]); xr[4 * i + 3] = fmaf(wgt, v.w, xr[4 * i + 3]);
                    }
                }
            }
            const float inv = 1.0f / den;
#pragma unroll
            for (int i = 0; i < 32; i++) xr[i] *= inv;
        }
    } else {
        if (loader) {
            const float4 *xp = reinterpret_cast<const float4 *>(P.x + (size_t)(blockIdx.x % P.x_reps) * P.x_stride + c * 32);
#pragma unroll
            for (int i = 0; i < 8; i++) { const float4 v = xp[i]; xr[4 * i] = v.x; xr[4 * i + 1] = v.y; xr[4 * i + 2] = v.z; xr[4 * i + 3] = v.w; }
        } else {
#pragma unroll
            for (int i = 0; i < 32; i++) xr[i] = 0.f;
        }
        float4 w4[8];
        if (P.xsrc == XS_RMSNORM && loader) {   // norm weights are requested together with x: one L2 round trip, not two
            const float4 *wp = reinterpret_cast<const float4 *>(P.norm_w + c * 32);
#pragma unroll
            for (int i = 0; i < 8; i++) w4[i] = wp[i];
        }
        if (threadIdx.x == 0) *(volatile int *)hold = p;   // input loads are on the wire: the producer may request this phase's weights
        if (P.xsrc == XS_RMSNORM) {  // RMSNormInto, go/quant.go:597-607: float64 sum of squares, fp32 x*inv*w
            // 32 squares per thread in four fp32 chains, everything across threads in float64 (the reference sums in float64;
            // the difference is far below one fp32 ulp of the resulting scale)
            float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
            if (active && rg == 0) {
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    q0 = fmaf(xr[4 * i], xr[4 * i], q0); q1 = fmaf(xr[4 * i + 1], xr[4 * i + 1], q1);
                    q2 = fmaf(xr[4 * i + 2], xr[4 * i + 2], q2); q3 = fmaf(xr[4 * i + 3], xr[4 * i + 3], q3);
                }
            }
            double ss = ((double)q0 + (double)q1) + ((double)q2 + (double)q3);
            if (threadIdx.x == 0 && ss >= 0.0) MG_TRACE(p, 6);   // x has landed
            ss = warp_sum_d(ss);
            if (lane == 0) ss_red[warp] = ss;
            mega_consumer_bar();
            if (warp == 0) {
                double v = lane < MG_CONSUMER_WARPS ? ss_red[lane] : 0.0;
                v = warp_sum_d(v);
                if (lane == 0) *inv_s = (float)(1.0 / sqrt(v / (double)P.cols + (double)A.eps));
            }
            mega_consumer_bar();
            if (threadIdx.x == 0) MG_TRACE(p, 7);                // scale known
            const float inv = *inv_s;
            if (loader) {
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    xr[4 * i] = xr[4 * i] * inv * w4[i].x; xr[4 * i + 1] = xr[4 * i + 1] * inv * w4[i].y;
                    xr[4 * i + 2] = xr[4 * i + 2] * inv * w4[i].z; xr[4 * i + 3] = xr[4 * i + 3] * inv * w4[i].w;
                }
            }
        }
    }
    if (share) {
        if (active && rg == 0) {
            float4 *xp = reinterpret_cast<float4 *>(xs + c * 32);
#pragma unroll
            for (int i = 0; i < 8; i++) xp[i] = make_float4(xr[4 * i], xr[4 * i + 1], xr[4 * i + 2], xr[4 * i + 3]);
        }
        mega_consumer_bar();
        if (active && rg != 0) {
            const float4 *xp = reinterpret_cast<const float4 *>(xs + c * 32);
#pragma unroll
            for (int i = 0; i < 8; i++) { const float4 v = xp[i]; xr[4 * i] = v.x; xr[4 * i + 1] = v.y; xr[4 * i + 2] = v.z; xr[4 * i + 3] = v.w; }
        }
    }
    float xoff = 0.f;
    if constexpr (TYPE == NL_Q4_0) {
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
        for (int i = 0; i < 8; i++) { s0 += xr[4 * i]; s1 += xr[4 * i + 1]; s2 += xr[4 * i + 2]; s3 += xr[4 * i + 3]; }
        xoff = -48.0f * ((s0 + s1) + (s2 + s3));
    }
    unsigned long long x2[16];
#pragma unroll
    for (int i = 0; i < 16; i++) x2[i] = pack2(xr[2 * i], xr[2 * i + 1]);

    if (threadIdx.x == 0) { MG_TRACE(p, 2); *(volatile int *)hold1 = 0; }
    // ---- stream the chunks ----
    // everything the inner loop needs lives in registers (the phase descriptor is in global memory)
    const int NS = A.stages, RG = P.RG, lg = P.upc_log2, n_part = NM << lg;
    const int tb = (rg < RG ? rg : RG - 1) * nb + (c < nb ? c : nb - 1);   // idle lanes shadow a valid block (result masked)
    const uint32_t mat_stride = P.q_chunk_bytes + P.d_chunk_bytes;
    const uint32_t unit_q = RG * nb * QB, unit_d = RG * nb * DB;
    const uint32_t base_q = tb * QB, base_d = P.q_chunk_bytes + tb * DB;
    const uint32_t slot_bytes = A.slot_bytes;
    const int red_idx = rg * wpr + wi;
    while (u < u_end) {
        const Chunk ch = next_chunk(P, u, u_end);
        const int slot = it % NS;
        mbar_wait(&full_bar[slot], (it / NS) & 1);
        const uint8_t *st = smem + (size_t)slot * slot_bytes;
        float part[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int pp = 0; pp < 2; pp++) {
            if (2 * pp < n_part && ((2 * pp) & ((1 << lg) - 1)) < ch.nunits) {  // uniform over the CTA; the two blocks below are straight-line
#pragma unroll
                for (int k = 0; k < 2; k++) {
                    const int pi = 2 * pp + k;
                    const int m = pi >> lg, r = pi & ((1 << lg) - 1);
                    const int mm = m < NM ? m : 0;
                    const uint8_t *mp = st + mm * mat_stride;
                    float v = block_dot2<TYPE>(mp + r * unit_q + base_q, x2) + xoff;
                    if (DB) v *= (TYPE == NL_Q4_0 ? 0.5f : 1.0f) * __half2float(*reinterpret_cast<const __half *>(mp + r * unit_d + base_d));
                    part[pi] = (active && m < NM && r * RG + rg < ch.rows) ? v : 0.f;
                }
            }
        }
        const float kk = warp_fold<4>(part, lane);
        if ((lane & 7) == 0 && rg < RG) red[slot][(lane >> 3) * 24 + red_idx] = kk;
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[slot]);
        it++;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// attention phase: item = (kv head, split).  RoPE on the group's q heads and on the new k (go/model.go:449-477, :530-539),
// optional bare QK-norm (:542-549), KV-cache write (:552-554), scores + softmax + V sum over this split's positions (:557-587).
// Results are un-normalised partials (flash-decoding); the O-projection phase combines them while loading its input.
struct AttnSmem {
    float q[MG_MAX_GROUP][64];
    float k[64], v[64];
    float sc[MG_MAX_GROUP][MG_ATT_CHUNK];
    float red[MG_CONSUMER_WARPS][MG_MAX_GROUP];
    float pv[3][MG_MAX_GROUP][64];
    float m_run[MG_MAX_GROUP], l_run[MG_MAX_GROUP], m_new[MG_MAX_GROUP], corr[MG_MAX_GROUP];
    int is_last;
};

template <int NT>
__device__ __forceinline__ void attn_item(const MegaAttn &at, int layer, int item, AttnSmem &S, int tid) {
    constexpr int HD = 64, HALF = 32;
    const int group = at.n_heads / at.n_kv_heads, kvd = at.n_kv_heads * HD;
    const int kvh = item / at.nsplit, sp = item % at.nsplit;
    const int pos = *at.pos, n = pos + 1;
    const int warp = tid >> 5, lane = tid & 31;
    // split [0, n) into nsplit chunks of whole 32-position groups (trailing items may be empty)
    int per = (n + at.nsplit - 1) / at.nsplit;
    per = (per + 31) / 32 * 32;
    const int t_begin = min(sp * per, n), t_end = min(t_begin + per, n);
    float *kc = at.kcache + (size_t)layer * at.seq_len * kvd, *vc = at.vcache + (size_t)layer * at.seq_len * kvd;

    // ---- RoPE (+QK-norm) of the group's q heads and of the new k; every item does it, only the owner of `pos` stores k/v
    const float *cs = at.cos_t + (size_t)pos * HALF, *sn = at.sin_t + (size_t)pos * HALF;
    for (int idx = tid; idx < (group + 1) * HALF; idx += NT) {
        const int hh = idx / HALF, i = idx % HALF;
        const bool isk = hh == group;
        const float *src = isk ? at.k + kvh * HD : at.q + (size_t)(kvh * group + hh) * HD;
        const float x0 = src[i], x1 = src[i + HALF], c = cs[i], s = sn[i];
        float r0, r1;
        if (!at.conj) { r0 = x0 * c - x1 * s; r1 = x0 * s + x1 * c; }
        else { r0 = x0 * c + x1 * s; r1 = -x0 * s + x1 * c; }
        float *dst = isk ? S.k : S.q[hh];
        dst[i] = r0; dst[i + HALF] = r1;
    }
    if (tid < HD) S.v[tid] = at.v[kvh * HD + tid];
    if (tid < group) { S.m_run[tid] = -INFINITY; S.l_run[tid] = 0.f; S.corr[tid] = 0.f; }
    cons_bar<NT>();
    if (at.qk_norm) {
        if (warp <= group) {  // RMSNormBare, go/quant.go:584-594
            float *vec = warp == group ? S.k : S.q[warp];
            double ss = 0.0;
            for (int i = lane; i < HD; i += 32) ss += (double)vec[i] * (double)vec[i];
            ss = warp_sum_d(ss);
            const float inv = (float)(1.0 / sqrt(ss / (double)HD + (double)at.eps));
            for (int i = lane; i < HD; i += 32) vec[i] *= inv;
        }
        cons_bar<NT>();
    }
    // exactly one item per kv head has a range that ends at n and is not empty: it owns position `pos` and stores the new row
    if (t_begin < n && t_end == n && tid < HD) {
        kc[(size_t)pos * kvd + kvh * HD + tid] = S.k[tid];
        vc[(size_t)pos * kvd + kvh * HD + tid] = S.v[tid];
    }

    // PV work split: thread = (position-interleaved part, head in group, dim)
    const int threads_per_part = group * HD;
    int nparts = NT / threads_per_part;
    if (nparts > 3) nparts = 3;
    const int my_part = tid / threads_per_part, my_h = (tid % threads_per_part) / HD, my_d = tid % HD;
    float acc = 0.f;

    const int sub = tid & 7, tg = tid >> 3;  // scores: 8 lanes per position, 96 positions per pass
    for (int c0 = t_begin; c0 < t_end; c0 += MG_ATT_CHUNK) {
        const int cn = min(c0 + MG_ATT_CHUNK, t_end) - c0;
        for (int tb = 0; tb < cn; tb += NT / 8) {
            const int tl = tb + tg, t = c0 + tl;
            const bool valid = tl < cn;
            float kreg[8];
            if (valid && t < pos) {
                const float4 *kp = reinterpret_cast<const float4 *>(kc + (size_t)t * kvd + kvh * HD + sub * 8);
                const float4 k0 = kp[0], k1 = kp[1];
                kreg[0] = k0.x; kreg[1] = k0.y; kreg[2] = k0.z; kreg[3] = k0.w; kreg[4] = k1.x; kreg[5] = k1.y; kreg[6] = k1.z; kreg[7] = k1.w;
            } else {
#pragma unroll
                for (int i = 0; i < 8; i++) kreg[i] = valid ? S.k[sub * 8 + i] : 0.f;
            }
            for (int hh = 0; hh < group; hh++) {
                float dot = 0.f;
#pragma unroll
                for (int i = 0; i < 8; i++) dot = fmaf(S.q[hh][sub * 8 + i], kreg[i], dot);
                dot += __shfl_xor_sync(0xffffffffu, dot, 1);
                dot += __shfl_xor_sync(0xffffffffu, dot, 2);
                dot += __shfl_xor_sync(0xffffffffu, dot, 4);
                if (sub == 0 && valid) S.sc[hh][tl] = dot * at.scale;
            }
        }
        cons_bar<NT>();
        if (warp < group) {  // running softmax statistics of head `warp` (flash-decoding form of go/quant.go:610-626)
            float mx = -INFINITY;
            for (int i = lane; i < cn; i += 32) mx = fmaxf(mx, S.sc[warp][i]);
            mx = warp_max(mx);
            const float m_old = S.m_run[warp], m_new = fmaxf(m_old, mx);
            float sum = 0.f;
            for (int i = lane; i < cn; i += 32) { const float e = expf(S.sc[warp][i] - m_new); S.sc[warp][i] = e; sum += e; }
            sum = warp_sum(sum);
            __syncwarp();
            if (lane == 0) {
                const float corr = (m_old == -INFINITY) ? 0.f : expf(m_old - m_new);
                S.corr[warp] = corr;
                S.l_run[warp] = S.l_run[warp] * corr + sum;
                S.m_run[warp] = m_new;
            }
        }
        cons_bar<NT>();
        if (my_part < nparts) {
            float a = acc * S.corr[my_h];
            for (int tl = my_part; tl < cn; tl += nparts) {
                const int t = c0 + tl;
                const float vv = (t < pos) ? vc[(size_t)t * kvd + kvh * HD + my_d] : S.v[my_d];
                a = fmaf(S.sc[my_h][tl], vv, a);
            }
            acc = a;
        }
        cons_bar<NT>();
    }
    if (my_part < nparts) S.pv[my_part][my_h][my_d] = acc;
    cons_bar<NT>();
    if (tid < threads_per_part) {
        float o = 0.f;
        for (int p = 0; p < nparts; p++) o += S.pv[p][my_h][my_d];
        const int h = kvh * group + my_h;
        at.part_acc[((size_t)(h * at.nsplit + sp)) * HD + my_d] = o;
        if (my_d == 0) {
            at.part_ml[(h * at.nsplit + sp) * 2] = S.m_run[my_h];
            at.part_ml[(h * at.nsplit + sp) * 2 + 1] = S.l_run[my_h];
        }
    }
    // The split that arrives last on this kv head folds all splits (fixed order => deterministic) and publishes the final
    // attention output, so the O-projection reads one plain vector instead of every CTA re-reading all partials.
    __threadfence();
    cons_bar<NT>();
    if (tid == 0) {
        unsigned int old;
        asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(at.split_cnt + layer * at.n_kv_heads + kvh) : "memory");
        S.is_last = (old == (unsigned)at.nsplit - 1u);
    }
    cons_bar<NT>();
    if (S.is_last && tid < threads_per_part) {
        const int h = kvh * group + my_h;
        float M = -INFINITY;
        for (int s = 0; s < at.nsplit; s++) M = fmaxf(M, __ldcg(at.part_ml + (h * at.nsplit + s) * 2));
        float den = 0.f, o = 0.f;
        for (int s = 0; s < at.nsplit; s++) {
            const float m = __ldcg(at.part_ml + (h * at.nsplit + s) * 2), l = __ldcg(at.part_ml + (h * at.nsplit + s) * 2 + 1);
            if (l > 0.f) {
                const float wgt = expf(m - M);
                den = fmaf(wgt, l, den);
                o = fmaf(wgt, __ldcg(at.part_acc + ((size_t)(h * at.nsplit + s)) * HD + my_d), o);
            }
        }
        o *= 1.0f / den;
        for (int r = 0; r < at.out_reps; r++) at.out[(size_t)r * at.out_stride + h * HD + my_d] = o;
    }
    cons_bar<NT>();  // S is reused by the next item of this CTA
}

// ---------------------------------------------------------------------------------------------------------------------
template <int TYPE>
__global__ void __launch_bounds__(MG_THREADS, 1) decode_mega_kernel(const MegaArgs A) {
    constexpr int QB = BlkBytes<TYPE>::Q, DB = BlkBytes<TYPE>::D;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t full_bar[MG_MAX_STAGES], empty_bar[MG_MAX_STAGES];
    __shared__ float red[MG_MAX_STAGES][MG_RED];
    __shared__ double ss_red[MG_CONSUMER_WARPS];
    __shared__ float inv_s;
    struct Pend { float *out; const float *bias; int phase, row0, rows, last, RG, wpr, upc, NM, epi, reps, stride; };
    __shared__ Pend pend[MG_MAX_STAGES];
    __shared__ AttnSmem att;
    __shared__ __align__(16) float xs_buf[MG_XS_FLOATS];
    __shared__ __align__(16) MegaPhase sph[2];
    __shared__ int hold1_flag;
    __shared__ int hold_flag;   // index of the newest phase whose input loads the math warps have issued (-1 at start)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int NS = A.stages, G = gridDim.x;
    if (tid == 0) {
        hold1_flag = 0;
        hold_flag = 0;   // phase 0 may stream right away (its input was produced by the previous kernel)
        for (int s = 0; s < NS; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], MG_CONSUMER_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == MG_CONSUMER_WARPS) {
        // ===================== producer + finisher warp =====================
        int *hold = &hold_flag;
        uint64_t policy;  // weights are read once per token: do not let them push activations / KV / norm weights out of L2
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
        auto finish = [&](int slot) {   // epilogue of the chunk that occupied `slot`; all 32 lanes; everything it needs is in smem
            const Pend q = pend[slot];
            for (int rr = lane; rr < q.rows; rr += 32) {
                const int r = rr / q.RG, rgi = rr - r * q.RG, row = q.row0 + rr;
                float v = 0.f, v2 = 0.f;
                for (int wi = 0; wi < q.wpr; wi++) {
                    v += red[slot][r * 24 + rgi * q.wpr + wi];
                    if (q.NM == 2) v2 += red[slot][(q.upc + r) * 24 + rgi * q.wpr + wi];
                }
                if (q.bias) v += q.bias[row];
                if (q.NM == 2) v = silu_f(v) * v2;                 // SiLU(gate)*up, go/model.go:604-606
                else if (q.epi == SEPI_RESID) v += __ldcg(q.out + row);     // X += W·x, go/model.go:592-594, :610-612
                for (int rep = 0; rep < q.reps; rep++) q.out[(size_t)rep * q.stride + row] = v;
            }
            __syncwarp();
            if (q.last && lane == 0) { MG_TRACE(q.phase, 5); phase_arrive(A.bar, q.phase); MG_TRACE(q.phase, 4); }  // release: cumulative over the warp
            __syncwarp();
        };
        // Issue and finish are decoupled: chunks are issued in program order whenever a ring slot is free and the math warps are
        // not in a phase prologue (a demand load issued while tens of KB of bulk data are inbound to this SM queues behind them
        // for microseconds, so the ring is not refilled between "phase p consumed" and "input of phase p+1 in registers");
        // consumed chunks are finished in order as soon as their empty-barrier completes.
        int issued = 0, finished = 0;
        int p = 0, u = 0, u_end = 0;
        bool have_phase = false, all_issued = false;
        while (!all_issued || finished < issued) {
            bool progress = false;
            if (finished < issued) {
                const int slot = finished % NS;
                int done = (lane == 0) ? (int)mbar_test(&empty_bar[slot], (finished / NS) & 1) : 0;
                done = __shfl_sync(0xffffffffu, done, 0);      // warp-uniform decision
                if (done) { finish(slot); finished++; progress = true; }
            }
            int ready = (lane == 0) ? *(volatile int *)hold : 0;
            ready = __shfl_sync(0xffffffffu, ready, 0);
            // `p` is the phase of the next chunk (or an earlier phase still to be skipped): chunks of phase q are only requested once
            // the math warps have put their input loads for phase q on the wire, so those loads never queue behind bulk data
            if (A.hold_mode != 2) ready = 0x7fffffff;
            int held1 = (lane == 0 && A.hold_mode == 1) ? *(volatile int *)&hold1_flag : 0;
            held1 = __shfl_sync(0xffffffffu, held1, 0);
            if (!all_issued && issued - finished < NS && !held1 && (have_phase ? p <= ready : true)) {
                while (!have_phase && p < A.n_phases) {   // advance to the next GEMV phase that has work for this CTA
                    const MegaPhase &P = A.phases[p];
                    if (P.kind == PH_GEMV) {
                        tile_band(P.total_units, blockIdx.x, G, u, u_end);
                        if (u < u_end) { have_phase = true; break; }
                        if (lane == 0) phase_arrive(A.bar, p);   // nothing of this phase lands here: arrive right away
                    }
                    p++;
                }
                if (!have_phase) { all_issued = true; continue; }
                if (p > ready) { __nanosleep(32); continue; }   // found the next phase with work, but its input is not requested yet
                const MegaPhase &P = A.phases[p];
                const Chunk ch = next_chunk(P, u, u_end);
                const int slot = issued % NS;
                if (lane == 0) {
                    const StreamSeg &sg = P.seg[ch.seg];
                    pend[slot] = Pend{sg.out, sg.bias, p, ch.row0, ch.rows, (int)(u == u_end), P.RG, P.nb_pad >> 5, P.upc, P.NM, P.epi, P.out_reps, P.out_stride};
                    const uint32_t mat_stride = P.q_chunk_bytes + P.d_chunk_bytes;
                    const uint32_t qb = (uint32_t)ch.rows * P.nb * QB, db = (uint32_t)ch.rows * P.nb * DB;
                    uint8_t *st = smem + (size_t)slot * A.slot_bytes;
                    mbar_expect_tx(&full_bar[slot], (qb + db) * P.NM);
                    bulk_g2s_hint(st, sg.qs + (size_t)ch.row0 * P.nb * QB, qb, &full_bar[slot], policy);
                    if (DB) bulk_g2s_hint(st + P.q_chunk_bytes, reinterpret_cast<const uint8_t *>(sg.d) + (size_t)ch.row0 * P.nb * DB, db, &full_bar[slot], policy);
                    if (P.NM == 2) {
                        bulk_g2s_hint(st + mat_stride, sg.qs2 + (size_t)ch.row0 * P.nb * QB, qb, &full_bar[slot], policy);
                        if (DB) bulk_g2s_hint(st + mat_stride + P.q_chunk_bytes, reinterpret_cast<const uint8_t *>(sg.d2) + (size_t)ch.row0 * P.nb * DB, db, &full_bar[slot], policy);
                    }
                }
                __syncwarp();
                issued++;
                if (u == u_end) { have_phase = false; p++; }
                progress = true;
            }
            if (!progress) {
                if (finished < issued) {   // nothing to issue right now: block on the oldest chunk and finish it the moment it is consumed
                    const int slot = finished % NS;
                    mbar_wait(&empty_bar[slot], (finished / NS) & 1);
                    finish(slot);
                    finished++;
                } else {
                    __nanosleep(32);
                }
            }
        }
        return;
    }

    // ===================== consumer warps =====================
    int it = 0;
    for (int p = 0; p < A.n_phases; p++) {
        // the (immutable) phase descriptor is fetched into shared memory while we wait for the previous phase to complete, so no
        // L2 round trip for it sits between the barrier and the first use
        if (warp == 1) {
            const uint32_t *src = reinterpret_cast<const uint32_t *>(&A.phases[p]);
            uint32_t *dst = reinterpret_cast<uint32_t *>(&sph[p & 1]);
            for (int i = lane; i < (int)(sizeof(MegaPhase) / 4); i += 32) dst[i] = __ldg(src + i);
        }
        const MegaPhase &P = sph[p & 1];
        if (tid == 0) { MG_TRACE(p, 0); if (p > 0) *(volatile int *)&hold1_flag = 1; }
        if (p > 0 && tid == 0) phase_wait(A.bar, p - 1, (unsigned)G);   // inputs of phase p are complete when every CTA arrived on p-1
        mega_consumer_bar();
        if (tid == 0) MG_TRACE(p, 1);
        if (P.kind == PH_ATTN) {
            if (tid == 0) { *(volatile int *)&hold_flag = p; *(volatile int *)&hold1_flag = 0; }
            const int n_items = A.at.n_kv_heads * A.at.nsplit;
            for (int item = blockIdx.x; item < n_items; item += G) attn_item<MG_CONSUMERS>(A.at, P.layer, item, att, tid);
            __threadfence();
            mega_consumer_bar();
            if (tid == 0) { MG_TRACE(p, 3); phase_arrive(A.bar, p); MG_TRACE(p, 4); }
            continue;
        }
        if (P.NM == 2) consume_phase<TYPE, 2>(A, P, smem, full_bar, empty_bar, red, ss_red, &inv_s, xs_buf, &hold_flag, &hold1_flag, it, warp, lane, p);
        else consume_phase<TYPE, 1>(A, P, smem, full_bar, empty_bar, red, ss_red, &inv_s, xs_buf, &hold_flag, &hold1_flag, it, warp, lane, p);
        if (tid == 0) MG_TRACE(p, 3);
    }
}

int launch_mega_q4_0(const MegaArgs &a, int grid, size_t smem, cudaStream_t st);
int launch_mega_q8_0(const MegaArgs &a, int grid, size_t smem, cudaStream_t st);
int launch_mega_f16(const MegaArgs &a, int grid, size_t smem, cudaStream_t st);

template <int TYPE>
int launch_mega_typed(const MegaArgs &a, int grid, size_t smem, cudaStream_t st) {
    static size_t configured = 0;  // largest dynamic smem opted into so far (static smem of the kernel comes on top)
    auto kern = decode_mega_kernel<TYPE>;
    if (smem > configured) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -2;
        configured = smem;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(MG_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative;  // all CTAs must be co-resident: they spin on each other
    at[0].val.cooperative = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, a) == cudaSuccess ? 0 : -2;
}

}  // namespace nl
