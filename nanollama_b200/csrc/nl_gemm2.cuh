// nl_gemm2.cuh — second-generation tcgen05 GEMM for the T-token form of matmulDispatch (go/model.go:361-386), sm_100a.
//
//   C_s[T, N_s] (=|+=) A[T, K] · W_s[N_s, K]^T   for up to three weight matrices W_s that share the input A (q|k|v, gate|up)
//
// Same arithmetic as nl_gemm.cuh (W dequantised on the fly into two bf16 planes, A given as two bf16 planes, three products accumulated
// in fp32 in TMEM); what changed is how the operands get to the tensor core.  The first kernel was bound by the way it fed itself: its
// A tiles were re-read from L2 in 16-byte pieces of 32-byte sectors (2x the bytes) once per 128 output columns, and only one K step
// of weights was in flight per SM.  Here:
//   * one CTA owns 256 output columns (tcgen05.mma M=128, N=256): half the A traffic per flop, and every sector that crosses the
//     L2 -> SM link is used whole (tile-order planes: 8 KB pieces; row-major planes: four lanes fetch one row's 64 bytes);
//   * K step 32 = one quant block per weight row: thread n dequantises row n's block, nothing else — 256 threads, 256 rows;
//   * separate rings: A tiles four deep (fetched two steps ahead), dequantised W two deep, raw quantised W in a shared-memory ring up
//     to 15 steps deep (cp.async, 60 KB in flight per SM: what an HBM-bound small batch needs);
//   * the activation planes come in TILE ORDER (plane_index, nl_common.cuh: their producers write a 128-row x 32-k tile as one
//     contiguous 8 KB piece, already in the UMMA core-matrix order) and the issuing thread fetches each tile with one cp.async.bulk
//     that completes on an mbarrier.  (a_tiled = 0: row-major planes, 256 threads x 8 cp.async per step -- their address registers are
//     held until the memory pipe takes each copy: 22 % of the producers' stall samples, 333 instead of 433 TFLOP/s at 2048 rows);
//   * two orientations.  WIDE (T > 128): activations are the M side, two 128-token tiles share every dequantised W tile (TMEM: 2 x 256
//     columns).  TALL (T <= 128): the WEIGHTS are the M side (two blocks of 128 rows) and the tokens are the N side, N = T rounded up
//     to 16 — a 16-sequence decode batch costs 1/8 of the tensor-core time of a 128-token tile instead of all of it, and the epilogue
//     writes 32 consecutive outputs of one token per warp.
//   * a ninth warp issues the MMAs.  The eight producer warps never wait for it: they arrive on a named barrier when their part of a
//     step is in shared memory and go on to the next step (blocking only when a ring slot is still being read); the issuing warp
//     syncs on that barrier, issues twelve tcgen05.mma with descriptors that cost one add each, commits to an mbarrier.  (First
//     version: thread 0 issued them after a __syncthreads -- ~35 dependent instructions per MMA in one thread, ~2000 cycles per step
//     in front of everybody's next step: 1900 cycles per 32-K step with a 16-token batch whose MMAs are almost free.)
// Rows beyond T or N are never zero-filled: a garbage A row only reaches its own (unstored) output row, a garbage W row its own column.
#pragma once
#include <cuda_bf16.h>

#include "nl_common.cuh"
#include "nl_gemm.cuh"   // descriptors, split2, umma_f16 / umma_commit, GEPI_*

namespace nl {

constexpr int G2_BN = 256, G2_BK = 32, G2_THREADS = 288;
constexpr int G2_A_TILE = 128 * G2_BK * 2;          // one bf16 plane of 128 activation rows x 32 k: 8 KB
constexpr int G2_W_TILE = G2_BN * G2_BK * 2;        // one bf16 plane of 256 weight rows x 32 k: 16 KB
constexpr int G2_A_SLOTS = 4, G2_W_SLOTS = 2;
constexpr int G2_A_LBO = 16 * 128, G2_W_LBO = 32 * 128, G2_SBO = 128;   // K-adjacent core matrices: 2048 | 4096 bytes apart
constexpr int G2_MAX_SEG = 3;

struct Gemm2Seg {
    const uint8_t *qs;     // planar quants (or raw F16 rows)
    const __half *d;       // block scales
    const float *bias;     // [N] or null
    float *c;              // [T][ldc]
    int N, ldc, epi;
    int tile_end;          // exclusive prefix sum of 256-column tiles over the segments
};
struct Gemm2Args {
    const __nv_bfloat16 *a_hi, *a_lo;   // [T][K] row-major
    Gemm2Seg seg[G2_MAX_SEG];
    int nseg, T, K;
    int a_tiled;   // the planes are in plane_index order (nl_common.cuh): the issuing warp fetches every 8 KB tile with ONE bulk copy
    // split K (tall only): grid.z CTAs share one tile, each takes `ksplit_steps` K steps and leaves its partial sums in
    // part[z][T][ldp] (column = 256 * tile + column in tile); gemm2_reduce_kernel adds them in z order and applies bias / residual
    int ksplit, ksplit_steps, ldp;
    float *part;
};

template <int TYPE> struct G2Raw {   // bytes of one weight row's K step in HBM, 16-byte chunks of it
    static constexpr int ROW = TYPE == NL_Q4_0 ? 16 : TYPE == NL_Q8_0 ? 32 : 64;
    static constexpr int CH = ROW / 16;
    static constexpr int STEP = 256 * ROW;
};
// MODE 0: tall (T <= 128).  MODE 1 / 2: wide with one / two 128-token tiles per CTA (one: when two would leave SMs without a CTA).
template <int TYPE, int MODE> struct G2Cfg {
    static constexpr bool WIDE = MODE != 0;
    static constexpr int MT = MODE == 2 ? 2 : 1;
    static constexpr int A_SLOT = MT * 2 * G2_A_TILE;                       // (hi, lo) per 128-row tile
    static constexpr int W_SLOT = 2 * G2_W_TILE;
    static constexpr int RAW_BYTES = MODE == 2 ? 32 * 1024 : 64 * 1024;
    static constexpr int RAW_SLOTS = RAW_BYTES / G2Raw<TYPE>::STEP;         // two tiles: 8 | 4 | 2, otherwise 16 | 8 | 4
    static constexpr int RD = RAW_SLOTS - 1;                                // raw prefetch distance in K steps
    static constexpr int A_OFF = 0, W_OFF = G2_A_SLOTS * A_SLOT, RAW_OFF = W_OFF + G2_W_SLOTS * W_SLOT;
    static constexpr int SMEM = RAW_OFF + RAW_BYTES;                        // two tiles: 224 KB, otherwise 192 KB
    static constexpr int TMEM_COLS = MODE == 2 ? 512 : 256;
};
constexpr int G2_PRODUCERS = 256;                    // warps 0..7: one weight row each; warp 8 issues the MMAs
// low word of a shared-memory matrix descriptor (the high word -- SBO, version -- is the same for every operand here)
constexpr uint32_t G2_DESC_HI = (uint32_t)(G2_SBO >> 4) | (1u << 14);
__device__ __forceinline__ uint32_t g2_desc_lo(uint32_t smem_addr, uint32_t lbo) { return ((smem_addr & 0x3FFFFu) >> 4) | ((lbo >> 4) << 16); }
__device__ __forceinline__ uint64_t g2_desc(uint32_t lo) {
    uint64_t d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(G2_DESC_HI));
    return d;
}
template <int ID> __device__ __forceinline__ void g2_bar_arrive() { asm volatile("bar.arrive %0, %1;" ::"n"(ID), "n"(G2_PRODUCERS + 32) : "memory"); }
template <int ID> __device__ __forceinline__ void g2_bar_sync() { asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(G2_PRODUCERS + 32) : "memory"); }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory"); }
__device__ __forceinline__ uint4 lds128u(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts128u(uint32_t a, const uint4 &v) { asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); }

// 8 fp32 weights -> one 16-byte chunk of each bf16 plane
__device__ __forceinline__ void g2_store8(const float *w, uint32_t hi_addr, uint32_t lo_addr) {
    uint4 vh, vl;
    split2(w[0], w[1], vh.x, vl.x); split2(w[2], w[3], vh.y, vl.y); split2(w[4], w[5], vh.z, vl.z); split2(w[6], w[7], vh.w, vl.w);
    sts128u(hi_addr, vh);
    sts128u(lo_addr, vl);
}

template <int TYPE, int MODE>
__global__ void __launch_bounds__(G2_THREADS, 1) gemm2_kernel(const Gemm2Args g) {
    using Cfg = G2Cfg<TYPE, MODE>;
    using Raw = G2Raw<TYPE>;
    constexpr bool WIDE = Cfg::WIDE;
    constexpr int MT = Cfg::MT, RD = Cfg::RD, RS = Cfg::RAW_SLOTS;
    static_assert(RD >= 2, "the raw ring has to run at least two K steps ahead (cp.async group accounting)");
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t mma_bar[2];
    __shared__ uint64_t a_full[G2_A_SLOTS];   // (tiled planes) the bulk copies of a step's activation tiles have landed
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // ---- which matrix, which 256 columns, which tokens
    int sidx = 0;
    while (sidx + 1 < g.nseg && (int)blockIdx.x >= g.seg[sidx].tile_end) sidx++;
    const Gemm2Seg &S = g.seg[sidx];
    const int n0 = ((int)blockIdx.x - (sidx ? g.seg[sidx - 1].tile_end : 0)) * G2_BN;
    const int m0 = WIDE ? (int)blockIdx.y * 128 * MT : 0;
    const int nb = g.K >> 5;
    const int ks0 = (!WIDE && g.ksplit > 1) ? (int)blockIdx.z * g.ksplit_steps : 0;                 // my K steps: [ks0, ks0 + ksteps)
    const int ksteps = (!WIDE && g.ksplit > 1) ? min(g.ksplit_steps, nb - ks0) : nb;
    const int Tp = WIDE ? 0 : ((g.T + 15) & ~15);      // tall: the MMA's N

    if (tid == 0) {
        mbar_init(&mma_bar[0], 1); mbar_init(&mma_bar[1], 1);
        for (int i = 0; i < G2_A_SLOTS; i++) mbar_init(&a_full[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(Cfg::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_c = tmem_base_s;
    const uint32_t sm = smem_u32(smem);

    if (warp == G2_PRODUCERS / 32) {
        // ===================== the issuing warp: one step behind the producers at most, never in their way =====================
        const uint32_t idesc = WIDE ? ((1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(G2_BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24))
                                    : ((1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(Tp >> 3) << 17) | ((uint32_t)(128 >> 4) << 24));
        const uint32_t a0 = g2_desc_lo(sm + Cfg::A_OFF, G2_A_LBO), w0 = g2_desc_lo(sm + Cfg::W_OFF, G2_W_LBO);
        // tiled planes: this thread also fetches the activation tiles, two steps ahead, one bulk copy per (128-row tile, plane); a tile
        // whose rows are all beyond T is skipped (the planes are allocated in whole tiles up to T)
        int n_mt = 0;
        for (int mt = 0; mt < MT; mt++) n_mt += (m0 + mt * 128 < g.T) ? 1 : 0;
        auto fetch_a = [&](int ks) {
            uint64_t *bar = &a_full[ks & (G2_A_SLOTS - 1)];
            uint8_t *slot = smem + Cfg::A_OFF + (size_t)(ks & (G2_A_SLOTS - 1)) * Cfg::A_SLOT;
            mbar_expect_tx(bar, (uint32_t)(n_mt * 2 * G2_A_TILE));
            for (int mt = 0; mt < n_mt; mt++) {
                const size_t e = ((size_t)((m0 >> 7) + mt) * (size_t)nb + (size_t)(ks0 + ks)) * (G2_A_TILE / 2);
                bulk_g2s(slot + (size_t)mt * (2 * G2_A_TILE), g.a_hi + e, G2_A_TILE, bar);
                bulk_g2s(slot + (size_t)mt * (2 * G2_A_TILE) + G2_A_TILE, g.a_lo + e, G2_A_TILE, bar);
            }
        };
        if (g.a_tiled && lane == 0) { fetch_a(0); if (ksteps > 1) fetch_a(1); }
        for (int ks = 0; ks < ksteps; ks++) {
            if (ks & 1) g2_bar_sync<2>(); else g2_bar_sync<1>();   // every producer's part of step ks is in shared memory (and fenced)
            if (lane == 0) {
                if (g.a_tiled) {
                    // (every producer has seen the commit of step ks - 2 before it arrived for step ks: this wait does not block, and it
                    // sits in front of this step's commit so that the barrier cannot be two phases ahead of the parity asked for)
                    if (ks + 2 < ksteps) { if (ks >= 2) mbar_wait(&mma_bar[ks & 1], ((ks >> 1) - 1) & 1); fetch_a(ks + 2); }
                    mbar_wait(&a_full[ks & (G2_A_SLOTS - 1)], (ks >> 2) & 1);
                }
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t al = a0 + (uint32_t)((ks & (G2_A_SLOTS - 1)) * (Cfg::A_SLOT >> 4));   // (descriptor words count 16-byte units)
                const uint32_t wl = w0 + (uint32_t)((ks & 1) * (Cfg::W_SLOT >> 4));
#pragma unroll
                for (int kk = 0; kk < G2_BK / 16; kk++) {   // one MMA consumes K = 16 = two core matrices along K
                    const uint32_t ka = (uint32_t)(kk * 2 * G2_A_LBO) >> 4, kw = (uint32_t)(kk * 2 * G2_W_LBO) >> 4;
                    if constexpr (WIDE) {
                        const uint64_t wh = g2_desc(wl + kw), wlo = g2_desc(wl + kw + (G2_W_TILE >> 4));
#pragma unroll
                        for (int mt = 0; mt < MT; mt++) {
                            const uint64_t ah = g2_desc(al + ka + (uint32_t)mt * ((2 * G2_A_TILE) >> 4)), alo = g2_desc(al + ka + (uint32_t)mt * ((2 * G2_A_TILE) >> 4) + (G2_A_TILE >> 4));
                            const uint32_t acc = tmem_c + (uint32_t)mt * G2_BN;   // accumulator mt: TMEM columns [256 mt, 256 mt + 256)
                            umma_f16(acc, ah, wh, idesc, (ks | kk) != 0);
                            umma_f16(acc, ah, wlo, idesc, 1);
                            umma_f16(acc, alo, wh, idesc, 1);
                        }
                    } else {
                        const uint64_t xh = g2_desc(al + ka), xl = g2_desc(al + ka + (G2_A_TILE >> 4));
#pragma unroll
                        for (int mb = 0; mb < 2; mb++) {   // weight rows [128 mb, 128 mb + 128) of the tile are the M side
                            const uint64_t wh = g2_desc(wl + kw + (uint32_t)mb * ((16 * G2_SBO) >> 4)), wlo = g2_desc(wl + kw + (uint32_t)mb * ((16 * G2_SBO) >> 4) + (G2_W_TILE >> 4));
                            const uint32_t acc = tmem_c + (uint32_t)mb * 128u;
                            umma_f16(acc, wh, xh, idesc, (ks | kk) != 0);
                            umma_f16(acc, wlo, xh, idesc, 1);
                            umma_f16(acc, wh, xl, idesc, 1);
                        }
                    }
                }
                umma_commit(&mma_bar[ks & 1]);   // arrives when every MMA issued so far has finished
            }
            __syncwarp();
        }
    } else {
    // ===================== producers =====================
    // ---- my weight row
    const int wn = tid;
    const bool w_ok = n0 + wn < S.N;
    const uint8_t *qrow = S.qs + (size_t)(w_ok ? n0 + wn : 0) * ((size_t)nb * Raw::ROW);   // (F16: K * 2 bytes per row = nb * 64)
    const __half *drow = TYPE == NL_F16 ? nullptr : S.d + (size_t)(w_ok ? n0 + wn : 0) * nb;
    const bool dvec = TYPE != NL_F16 && (nb & 7) == 0;                                     // a row's scales in 16-byte pieces (8 K steps)
    uint4 dcur = make_uint4(0u, 0u, 0u, 0u), dnext = make_uint4(0u, 0u, 0u, 0u);
    float dsc_next = 0.f;
    if (TYPE != NL_F16 && w_ok) {
        if (dvec) { dcur = __ldg(reinterpret_cast<const uint4 *>(drow) + (ks0 >> 3)); if ((ks0 >> 3) + 1 < (nb >> 3)) dnext = __ldg(reinterpret_cast<const uint4 *>(drow) + (ks0 >> 3) + 1); }
        else dsc_next = __half2float(drow[ks0]);
    }

    // ---- producers of one K step
    // (ks below: step index inside this CTA's K range -- ring slots; ks0 + ks: the step inside the matrix -- addresses)
    auto issue_raw = [&](int ks) {     // my row's quantised block -> raw ring, [chunk][row][16 B]: conflict-free to write and to read back
        if (!w_ok) return;
        const uint32_t dst = sm + Cfg::RAW_OFF + (uint32_t)(ks % RS) * Raw::STEP + (uint32_t)wn * 16u;
        const uint8_t *src = qrow + (size_t)(ks0 + ks) * Raw::ROW;
#pragma unroll
        for (int c = 0; c < Raw::CH; c++) cp_async16(dst + (uint32_t)c * 4096u, src + 16 * c);
    };
    auto issue_a = [&](int ks) {       // activation tiles of the step: lanes 4r..4r+3 fetch the 64 bytes of row r (both planes)
        const uint32_t slot = sm + Cfg::A_OFF + (uint32_t)(ks & (G2_A_SLOTS - 1)) * Cfg::A_SLOT;
#pragma unroll
        for (int j = 0; j < 2 * MT; j++) {
            const int q = tid + G2_PRODUCERS * j, kc = q & 3, r = q >> 2;   // r < 128 * MT
            if (m0 + r < g.T) {
                const size_t off = (size_t)(m0 + r) * g.K + (size_t)(ks0 + ks) * G2_BK + kc * 8;
                const uint32_t dh = slot + (uint32_t)(r >> 7) * (2 * G2_A_TILE) + (uint32_t)kc * G2_A_LBO + (uint32_t)((r & 127) >> 3) * G2_SBO + (uint32_t)(r & 7) * 16u;
                cp_async16(dh, g.a_hi + off);
                cp_async16(dh + G2_A_TILE, g.a_lo + off);
            }
        }
    };

    // ---- prologue: the raw ring RD steps deep, the A ring two steps deep; one cp.async group per (virtual) iteration
    for (int v = -RD; v < 0; v++) {
        if (v + RD < ksteps) issue_raw(v + RD);
        if (!g.a_tiled && v + 2 >= 0 && v + 2 < ksteps) issue_a(v + 2);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }

    for (int ks = 0; ks < ksteps; ks++) {
        // stage reuse: the MMAs of step ks - 2 read W slot ks & 1 and A slot (ks + 2) & 3
        if (ks >= 2) mbar_wait(&mma_bar[ks & 1], ((ks >> 1) - 1) & 1);
        if (!g.a_tiled && ks + 2 < ksteps) issue_a(ks + 2);
        // (the raw slot of step ks + RD is the one step ks - 1 was read from: RS = RD + 1 slots)
        if (ks + RD < ksteps) issue_raw(ks + RD);
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 2;" ::: "memory");   // everything up to the group of iteration ks - 2: A(ks), raw(ks)

        // ---- my row's block of step ks: raw ring -> registers -> two bf16 planes of the W tile
        float dsc = 0.f;
        if (TYPE != NL_F16) {
            if (dvec) {
                const int kg = ks0 + ks;
                if (ks && (kg & 7) == 0) { dcur = dnext; if (w_ok && (kg >> 3) + 1 < (nb >> 3)) dnext = __ldg(reinterpret_cast<const uint4 *>(drow) + (kg >> 3) + 1); }
                const int e = kg & 7;
                const uint32_t wsel = e < 4 ? (e < 2 ? dcur.x : dcur.y) : (e < 6 ? dcur.z : dcur.w);
                dsc = __half2float(__ushort_as_half((unsigned short)((e & 1) ? (wsel >> 16) : (wsel & 0xFFFFu))));
            } else {
                dsc = dsc_next;
                if (w_ok && ks + 1 < ksteps) dsc_next = __half2float(drow[ks0 + ks + 1]);
            }
        }
        {
            const uint32_t raw = sm + Cfg::RAW_OFF + (uint32_t)(ks % RS) * Raw::STEP + (uint32_t)wn * 16u;
            const uint32_t wt = sm + Cfg::W_OFF + (uint32_t)(ks & 1) * Cfg::W_SLOT + (uint32_t)(wn >> 3) * G2_SBO + (uint32_t)(wn & 7) * 16u;   // + kc * G2_W_LBO
            float w[8];
            if constexpr (TYPE == NL_Q4_0) {
                const uint4 q = lds128u(raw);
                const uint32_t ws[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int hn = 0; hn < 2; hn++) {          // low nibbles: weights 0..15, high nibbles: 16..31
#pragma unroll
                    for (int c = 0; c < 2; c++) {         // chunk kc = 2 * hn + c takes bytes 8c .. 8c+7
#pragma unroll
                        for (int i = 0; i < 2; i++) {
                            const uint32_t n4 = (hn ? (ws[2 * c + i] >> 4) : ws[2 * c + i]) & 0x0F0F0F0Fu;
#pragma unroll
                            for (int k = 0; k < 4; k++) w[4 * i + k] = (u8_to_f32_magic(n4, k) - 8388616.0f) * dsc;   // exact int -> fp32: 2^23 + n
                        }
                        g2_store8(w, wt + (uint32_t)(2 * hn + c) * G2_W_LBO, wt + (uint32_t)(2 * hn + c) * G2_W_LBO + G2_W_TILE);
                    }
                }
            } else if constexpr (TYPE == NL_Q8_0) {
#pragma unroll
                for (int hb = 0; hb < 2; hb++) {
                    const uint4 q = lds128u(raw + (uint32_t)hb * 4096u);
                    const uint32_t ws[4] = {q.x ^ 0x80808080u, q.y ^ 0x80808080u, q.z ^ 0x80808080u, q.w ^ 0x80808080u};
#pragma unroll
                    for (int c = 0; c < 2; c++) {
#pragma unroll
                        for (int i = 0; i < 2; i++)
#pragma unroll
                            for (int k = 0; k < 4; k++) w[4 * i + k] = (u8_to_f32_magic(ws[2 * c + i], k) - 8388736.0f) * dsc;
                        g2_store8(w, wt + (uint32_t)(2 * hb + c) * G2_W_LBO, wt + (uint32_t)(2 * hb + c) * G2_W_LBO + G2_W_TILE);
                    }
                }
            } else {
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const uint4 q = lds128u(raw + (uint32_t)c * 4096u);
                    const uint32_t ws[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&ws[i]));
                        w[2 * i] = f.x; w[2 * i + 1] = f.y;
                    }
                    g2_store8(w, wt + (uint32_t)c * G2_W_LBO, wt + (uint32_t)c * G2_W_LBO + G2_W_TILE);
                }
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes (cp.async, st.shared) -> visible to the tensor core
        if (ks & 1) g2_bar_arrive<2>(); else g2_bar_arrive<1>();       // (two barriers in turn: a producer can be a step ahead of the slowest one)
    }
    }   // producers

    // ---- epilogue: the last commit covers all earlier MMAs
    {
        const int last = ksteps - 1;
        mbar_wait(&mma_bar[last & 1], (last >> 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int lg = warp & 3, hf = warp >> 2;   // TMEM lane group of this warp; column half (wide) | weight block (tall)
        if (warp >= G2_PRODUCERS / 32) {
            // (the issuing warp has no rows to store)
        } else if constexpr (WIDE) {
#pragma unroll 1
            for (int mc = 0; mc < 4 * MT; mc++) {
                const int mt = mc >> 2, cc = mc & 3;
                const int row = m0 + mt * 128 + lg * 32 + lane;
                const int col0 = hf * 128 + cc * 32;
                uint32_t r[32];
                const uint32_t taddr = tmem_c + ((uint32_t)(lg * 32) << 16) + (uint32_t)(mt * G2_BN + col0);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,"
                    "%28,%29,%30,%31}, [%32];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
                      "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
                      "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (row < g.T && n0 + col0 < S.N) {
                    float *cp = S.c + (size_t)row * S.ldc + n0 + col0;
                    if (n0 + col0 + 32 <= S.N && (S.ldc & 3) == 0) {   // whole 128-byte run of this row: vector stores
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            float4 v = make_float4(__uint_as_float(r[i]), __uint_as_float(r[i + 1]), __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
                            if (S.bias) { const float *b = S.bias + n0 + col0 + i; v.x += __ldg(b); v.y += __ldg(b + 1); v.z += __ldg(b + 2); v.w += __ldg(b + 3); }
                            if (S.epi == GEPI_RESID) { const float4 o = *reinterpret_cast<const float4 *>(cp + i); v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
                            *reinterpret_cast<float4 *>(cp + i) = v;
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; i++) {
                            const int n = n0 + col0 + i;
                            if (n < S.N) {
                                float v = __uint_as_float(r[i]);
                                if (S.bias) v += S.bias[n];
                                if (S.epi == GEPI_RESID) v += cp[i];
                                cp[i] = v;
                            }
                        }
                    }
                }
            }
        } else {
            // D is [weight row][token]: lane = weight row, column = token; a warp stores 32 consecutive outputs of one token at a time
            const int n = n0 + hf * 128 + lg * 32 + lane;
            const bool split = g.ksplit > 1;
            const bool ok = n < S.N;
            const float bv = (ok && S.bias && !split) ? S.bias[n] : 0.f;
            float *pz = split ? g.part + (size_t)blockIdx.z * g.T * g.ldp + (size_t)blockIdx.x * G2_BN + hf * 128 + lg * 32 + lane : nullptr;
#pragma unroll 1
            for (int t0 = 0; t0 < Tp; t0 += 16) {
                uint32_t r[16];
                const uint32_t taddr = tmem_c + ((uint32_t)(lg * 32) << 16) + (uint32_t)(hf * 128 + t0);
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                             : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
                               "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                             : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (ok) {
#pragma unroll
                    for (int j = 0; j < 16; j++) {
                        const int t = t0 + j;
                        if (t < g.T) {
                            if (split) pz[(size_t)t * g.ldp] = __uint_as_float(r[j]);
                            else {
                                float *cp = S.c + (size_t)t * S.ldc + n;
                                float v = __uint_as_float(r[j]) + bv;
                                if (S.epi == GEPI_RESID) v += *cp;
                                *cp = v;
                            }
                        }
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_c), "n"(Cfg::TMEM_COLS) : "memory");
}

// split K, second half: out[t][n] = sum over z (in z order: deterministic) of part[z][t][column] (+ bias, + residual)
static __global__ void gemm2_reduce_kernel(const Gemm2Args g) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x, t = blockIdx.y;
    if (col >= g.ldp) return;
    const int tile = col / G2_BN;
    int sidx = 0;
    while (sidx + 1 < g.nseg && tile >= g.seg[sidx].tile_end) sidx++;
    const Gemm2Seg &S = g.seg[sidx];
    const int n = col - (sidx ? g.seg[sidx - 1].tile_end : 0) * G2_BN;
    if (n >= S.N) return;
    const float *p = g.part + (size_t)t * g.ldp + col;
    float v = 0.f;
    for (int z = 0; z < g.ksplit; z++) v += p[(size_t)z * g.T * g.ldp];
    if (S.bias) v += S.bias[n];
    float *cp = S.c + (size_t)t * S.ldc + n;
    if (S.epi == GEPI_RESID) v += *cp;
    *cp = v;
}

template <int TYPE, int MODE>
static int launch_gemm2_t(const Gemm2Args &g, cudaStream_t st) {
    using Cfg = G2Cfg<TYPE, MODE>;
    static bool configured = false;
    auto kern = gemm2_kernel<TYPE, MODE>;
    if (!configured) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM) != cudaSuccess) return -2;
        configured = true;
    }
    const bool split = MODE == 0 && g.ksplit > 1;
    dim3 grid(g.seg[g.nseg - 1].tile_end, MODE == 0 ? 1 : (g.T + 128 * Cfg::MT - 1) / (128 * Cfg::MT), split ? g.ksplit : 1);
    kern<<<grid, G2_THREADS, Cfg::SMEM, st>>>(g);
    if (split) gemm2_reduce_kernel<<<dim3((g.ldp + 255) / 256, g.T), 256, 0, st>>>(g);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
template <int TYPE>
int launch_gemm2_typed(const Gemm2Args &g, cudaStream_t st) {
    if (g.T <= 128) return launch_gemm2_t<TYPE, 0>(g, st);
    // two 128-token tiles per CTA share every dequantised W tile -- unless that leaves SMs without a CTA: then one tile each
    const int tiles = g.seg[g.nseg - 1].tile_end;
    return tiles * ((g.T + 255) / 256) >= 148 ? launch_gemm2_t<TYPE, 2>(g, st) : launch_gemm2_t<TYPE, 1>(g, st);
}
int launch_gemm2_q4_0(const Gemm2Args &g, cudaStream_t st);
int launch_gemm2_q8_0(const Gemm2Args &g, cudaStream_t st);
int launch_gemm2_f16(const Gemm2Args &g, cudaStream_t st);   // (wide: one 128-token tile per CTA)

}  // namespace nl
