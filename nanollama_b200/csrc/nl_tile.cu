// nl_tile.cu — kernels of the tiled tensor-core decode path (see nl_tile.cuh).
#include "nl_tile.cuh"

namespace nl {

// ---- layout builder: planar (qs, d) -> tiles.  Row group R of the source lands at tile row group rg_off + R * rg_stride
// (gate/up interleave: stride 2, offsets 0 / 1; q,k,v concatenation: stride 1, running offsets).
__global__ void tile_q4_0_kernel(const uint4 *__restrict__ qs, const __half *__restrict__ d, int rows, int nb, uint8_t *__restrict__ tiles,
                                 int nbg, int rg_off, int rg_stride, int n_rg_src) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = (int)(gid & 31);
    const long long tile = gid >> 5;
    if (tile >= (long long)n_rg_src * nbg) return;
    const int R = (int)(tile / nbg), B = (int)(tile % nbg);
    const int g = lane >> 2, t = lane & 3;
    const int blk = 4 * B + t;
    uint8_t *dst = tiles + ((size_t)(rg_off + R * rg_stride) * nbg + B) * TL_TILE;
    uint4 q0 = make_uint4(0, 0, 0, 0), q1 = q0;
    unsigned short d0 = 0, d1 = 0;
    const int r0 = 16 * R + g, r1 = r0 + 8;
    if (blk < nb) {
        if (r0 < rows) { q0 = qs[(size_t)r0 * nb + blk]; d0 = __half_as_ushort(d[(size_t)r0 * nb + blk]); }
        if (r1 < rows) { q1 = qs[(size_t)r1 * nb + blk]; d1 = __half_as_ushort(d[(size_t)r1 * nb + blk]); }
    }
    reinterpret_cast<uint4 *>(dst)[lane] = q0;
    reinterpret_cast<uint4 *>(dst + 512)[lane] = q1;
    reinterpret_cast<uint32_t *>(dst + 1024)[lane] = (uint32_t)d0 | ((uint32_t)d1 << 16);
}

// ---- device helpers ----
__device__ __forceinline__ void mma_f16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
    const __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<const uint32_t *>(&h);
}
template <int NT> __device__ __forceinline__ void tl_bar() { asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory"); }
__device__ __forceinline__ void band_of(int units, int b, int G, int &u0, int &u1) {
    u0 = (int)(((long long)units * b) / G);
    u1 = (int)(((long long)units * (b + 1)) / G);
}
// tile index inside a band -> row group (magic = ceil(2^32 / nbg); exact for r < 2^32 / nbg)
__device__ __forceinline__ int rg_of(int r, int nbg, unsigned int magic) { return nbg == 1 ? r : (int)__umulhi((unsigned)r, magic); }
template <typename T> __device__ __forceinline__ T *ldg_ptr(T *const *p) {
    return reinterpret_cast<T *>(__ldg(reinterpret_cast<const unsigned long long *>(p)));
}
#define TL_TRACE(p, k) do { if (A.trace) A.trace[((size_t)blockIdx.x * A.n_phases + (p)) * 8 + (k)] = gtime(); } while (0)

// One tile: 16 rows x 4 blocks.  acc0 / acc1: this lane's running sums for (row g, block column t) and (row g+8, t).
__device__ __forceinline__ void tile_dot(const uint8_t *tp, const uint8_t *xfrag_b, const float corr_v, bool xact, int lane, uint32_t (&xb)[16],
                                         float &acc0, float &acc1) {
    const uint4 wa4 = *reinterpret_cast<const uint4 *>(tp + lane * 16);
    const uint4 wb4 = *reinterpret_cast<const uint4 *>(tp + 512 + lane * 16);
    const uint32_t dd = *reinterpret_cast<const uint32_t *>(tp + 1024 + lane * 4);
    if (xact) {   // only the lane that owns (block column, hi|lo) of B holds data; every other lane keeps zeros
        const uint4 *xp = reinterpret_cast<const uint4 *>(xfrag_b);
#pragma unroll
        for (int i = 0; i < 4; i++) { const uint4 v = xp[i]; xb[4 * i] = v.x; xb[4 * i + 1] = v.y; xb[4 * i + 2] = v.z; xb[4 * i + 3] = v.w; }
    }
    const uint32_t wa[4] = {wa4.x, wa4.y, wa4.z, wa4.w}, wb[4] = {wb4.x, wb4.y, wb4.z, wb4.w};
    float c[4] = {0.f, 0.f, 0.f, 0.f}, e[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const uint32_t a8 = wa[i] >> 8, b8 = wb[i] >> 8;
        // elements 4i..4i+3 (low nibbles): n * 2^-24 as fp16 subnormals; B carries 16 * x * S
        mma_f16(c, wa[i] & 0x000F000Fu, wb[i] & 0x000F000Fu, a8 & 0x000F000Fu, b8 & 0x000F000Fu, xb[4 * i], xb[4 * i + 1]);
        // elements 16+4i..16+4i+3 (high nibbles): 16n * 2^-24; B carries x * S
        mma_f16(e, wa[i] & 0x00F000F0u, wb[i] & 0x00F000F0u, a8 & 0x00F000F0u, b8 & 0x00F000F0u, xb[4 * i + 2], xb[4 * i + 3]);
    }
    const float2 df = __half22float2(*reinterpret_cast<const __half2 *>(&dd));
    const float v0 = ((c[0] + e[0]) + (c[1] + e[1])) + corr_v;   // (hi + lo columns) - 8 * sum(x) of the block
    const float v1 = ((c[2] + e[2]) + (c[3] + e[3])) + corr_v;
    acc0 = fmaf(df.x, v0, acc0);
    acc1 = fmaf(df.y, v1, acc1);
}

struct TlShared {
    uint64_t full_bar[TL_SLOTS], empty_bar[TL_SLOTS], free_bar[TL_SLOTS];
    float red[TL_SLOTS][TL_CW][2][16];
    double ss_red[TL_CW];
    float mx_red[TL_CW];
    float post_scale[2];
    TilePhase ph[2];
};

__global__ void __launch_bounds__(TL_THREADS, 1) decode_tiled_kernel(const TileArgs A) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ TlShared sh;
    uint8_t *ring = smem;
    uint8_t *xfrag = smem + (size_t)TL_SLOTS * TL_SLOT_BYTES;
    float *corr = reinterpret_cast<float *>(xfrag + TL_XFRAG_BYTES);
    AttnSmem &att = *reinterpret_cast<AttnSmem *>(xfrag);   // the attention phase has no GEMV input: same bytes
    static_assert(sizeof(AttnSmem) <= TL_XFRAG_BYTES, "attention scratch must fit the fragment buffer");

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int G = gridDim.x;
    if (tid == 0) {
        for (int s = 0; s < TL_SLOTS; s++) { mbar_init(&sh.full_bar[s], 1); mbar_init(&sh.empty_bar[s], TL_CW); mbar_init(&sh.free_bar[s], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == TL_CW) {
        // ===================== copy warp: streams this CTA's band of every GEMV phase, in phase order =====================
        if (lane != 0) return;
        uint64_t policy;   // weights are read once per token: keep them from evicting activations / KV / norm weights out of L2
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
        int it = 0;
        for (int p = 0; p < A.n_phases; p++) {
            const TilePhase *P = A.phases + p;
            if (__ldg(&P->kind) != PH_GEMV) continue;
            const int nbg = __ldg(&P->nbg), urg = __ldg(&P->unit_rg);
            int u0, u1;
            band_of(__ldg(&P->n_rg) / urg, blockIdx.x, G, u0, u1);
            const int band = (u1 - u0) * urg * nbg;
            const uint8_t *src = ldg_ptr(&P->tiles) + (size_t)u0 * urg * nbg * TL_TILE;
            for (int c0 = 0; c0 < band; c0 += TL_TS, it++) {
                const int slot = it % TL_SLOTS;
                if (it >= TL_SLOTS) mbar_wait(&sh.free_bar[slot], ((it / TL_SLOTS) - 1) & 1);
                const uint32_t bytes = (uint32_t)min(TL_TS, band - c0) * TL_TILE;
                mbar_expect_tx(&sh.full_bar[slot], bytes);
                bulk_g2s_hint(ring + (size_t)slot * TL_SLOT_BYTES, src + (size_t)c0 * TL_TILE, bytes, &sh.full_bar[slot], policy);
            }
        }
        return;
    }

    if (warp == TL_CW + 1) {
        // ===================== finishing warp =====================
        const int row = lane & 15, half = lane >> 4;
        int it = 0;
        for (int p = 0; p < A.n_phases; p++) {
            const TilePhase *P = A.phases + p;
            if (__ldg(&P->kind) != PH_GEMV) continue;
            const int nbg = __ldg(&P->nbg), urg = __ldg(&P->unit_rg), epi = __ldg(&P->epi), rows = __ldg(&P->rows);
            const unsigned int magic = __ldg(&P->nbg_magic);
            const float *bias = ldg_ptr(&P->bias);
            float *out = ldg_ptr(&P->out);
            int u0, u1;
            band_of(__ldg(&P->n_rg) / urg, blockIdx.x, G, u0, u1);
            const int band = (u1 - u0) * urg * nbg;
            const int rg0 = u0 * urg;
            float racc = 0.f, gate = 0.f, post = 1.f;
            for (int c0 = 0; c0 < band; c0 += TL_TS, it++) {
                const int slot = it % TL_SLOTS;
                const int c1 = min(c0 + TL_TS, band);
                const int q_first = rg_of(c0, nbg, magic), q_last = rg_of(c1 - 1, nbg, magic);
                // the residual of a row group that completes in this slot is fetched before we block on the math warps (not for the first
                // slot of a phase: only a consumed slot proves that this CTA is past the grid barrier that orders the residual's writers)
                float resid = 0.f;
                int q_done = -1;
                if (epi == TEPI_RESID && c0 > 0) {
                    for (int q = q_first; q <= q_last; q++) if ((q + 1) * nbg <= c1) { q_done = q; break; }
                    if (q_done >= 0 && half == 0) { const int r = (rg0 + q_done) * 16 + row; if (r < rows) resid = __ldcg(out + r); }
                }
                mbar_wait(&sh.empty_bar[slot], (it / TL_SLOTS) & 1);
                if (c0 == 0) post = sh.post_scale[p & 1];
                for (int q = q_first; q <= q_last; q++) {
                    const int a = max(q * nbg, c0) - c0, b = min((q + 1) * nbg, c1) - c0;   // tiles [a, b) of this slot belong to row group q
                    const int wa = a >> 1, wb = (b - 1) >> 1;
                    float s = 0.f;
                    for (int w = wa + half; w <= wb; w += 2) {
                        const int e = (rg_of(c0 + 2 * w, nbg, magic) == q) ? 0 : 1;
                        s += sh.red[slot][w][e][row];
                    }
                    s += __shfl_xor_sync(0xffffffffu, s, 16);
                    racc += s;
                    if ((q + 1) * nbg <= c1) {   // last tile of the row group is in this slot: publish its 16 rows
                        const int rg = rg0 + q;
                        float v = racc * post;
                        racc = 0.f;
                        if (epi == TEPI_SWIGLU) {
                            if ((rg & 1) == 0) gate = v;
                            else {
                                const int r = (rg >> 1) * 16 + row;
                                if (half == 0 && r < rows) out[r] = silu_f(gate) * v;          // SiLU(gate)*up, go/model.go:604-606
                            }
                        } else {
                            const int r = rg * 16 + row;
                            if (half == 0 && r < rows) {
                                if (bias) v += __ldg(bias + r);
                                if (epi == TEPI_RESID) v += (q == q_done) ? resid : __ldcg(out + r);   // X += W.x, go/model.go:592-594, :610-612
                                out[r] = v;
                            }
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&sh.free_bar[slot]);
            }
            __syncwarp();
            if (lane == 0) { TL_TRACE(p, 4); phase_arrive(A.bar, p); }   // release: cumulative over the warp's stores
            __syncwarp();
        }
        return;
    }

    // ===================== math warps =====================
    const int g = lane >> 2, t = lane & 3;
    const bool xact = (t == (g >> 1));     // lane that holds B column g (block column g>>1, hi|lo = g&1)
    uint32_t xb[16];
#pragma unroll
    for (int i = 0; i < 16; i++) xb[i] = 0u;
    int it = 0;
    if (warp == 1) {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(&A.phases[0]);
        uint32_t *dst = reinterpret_cast<uint32_t *>(&sh.ph[0]);
        for (int i = lane; i < (int)(sizeof(TilePhase) / 4); i += 32) dst[i] = __ldg(src + i);
    }
    for (int p = 0; p < A.n_phases; p++) {
        if (warp == 1 && p + 1 < A.n_phases) {   // next descriptor while this phase runs: no L2 round trip after the barrier
            const uint32_t *src = reinterpret_cast<const uint32_t *>(&A.phases[p + 1]);
            uint32_t *dst = reinterpret_cast<uint32_t *>(&sh.ph[(p + 1) & 1]);
            for (int i = lane; i < (int)(sizeof(TilePhase) / 4); i += 32) dst[i] = __ldg(src + i);
        }
        if (p == 0) tl_bar<TL_CONSUMERS>();
        // every field of the descriptor is copied out before the last barrier of the prologue: warp 1 recycles the slot for phase
        // p + 2 as soon as it gets there
        const TilePhase &P = sh.ph[p & 1];
        const int kind = P.kind, layer = P.layer, nbg = P.nbg, cols = P.cols;
        const unsigned int magic = P.nbg_magic;
        const float *px = P.x, *pnw = P.norm_w;
        int u0, u1;
        band_of(P.n_rg / P.unit_rg, blockIdx.x, G, u0, u1);
        const int band = (kind == PH_GEMV) ? (u1 - u0) * P.unit_rg * nbg : 0;
        if (kind == PH_ATTN) {
            if (tid == 0) { TL_TRACE(p, 0); phase_wait(A.bar, p - 1, (unsigned)G); TL_TRACE(p, 1); }
            tl_bar<TL_CONSUMERS>();
            const int n_items = A.at.n_kv_heads * A.at.nsplit;
            for (int item = blockIdx.x; item < n_items; item += G) attn_item<TL_CONSUMERS>(A.at, layer, item, att, tid);
            __threadfence();
            tl_bar<TL_CONSUMERS>();
            if (tid == 0) { TL_TRACE(p, 3); phase_arrive(A.bar, p); TL_TRACE(p, 4); }
            continue;
        }

        // ---- prologue: phase input -> fp16 hi/lo B fragments in shared memory ----
        const int nitem = cols >> 3, nitem_pad = nbg * 16;
        const bool normed = pnw != nullptr;
        float wv[TL_MAX_ITEMS][8];
        if (normed && band > 0) {   // static data: requested before we wait for the producers of x
#pragma unroll
            for (int r = 0; r < TL_MAX_ITEMS; r++) {
                const int q = tid + r * TL_CONSUMERS;
                if (q < nitem) {
                    const float4 a = __ldg(reinterpret_cast<const float4 *>(pnw) + 2 * q), b = __ldg(reinterpret_cast<const float4 *>(pnw) + 2 * q + 1);
                    wv[r][0] = a.x; wv[r][1] = a.y; wv[r][2] = a.z; wv[r][3] = a.w; wv[r][4] = b.x; wv[r][5] = b.y; wv[r][6] = b.z; wv[r][7] = b.w;
                }
            }
        }
        if (p > 0) {
            if (tid == 0) { TL_TRACE(p, 0); phase_wait(A.bar, p - 1, (unsigned)G); TL_TRACE(p, 1); }
            tl_bar<TL_CONSUMERS>();
        }
        if (band == 0) continue;   // nothing of this matrix lands here (the finishing warp has arrived for us)
        float xv[TL_MAX_ITEMS][8];
        double ss = 0.0;
        float mx = 0.f;
#pragma unroll
        for (int r = 0; r < TL_MAX_ITEMS; r++) {
            const int q = tid + r * TL_CONSUMERS;
            if (q < nitem) {
                const float4 a = __ldcg(reinterpret_cast<const float4 *>(px) + 2 * q), b = __ldcg(reinterpret_cast<const float4 *>(px) + 2 * q + 1);
                xv[r][0] = a.x; xv[r][1] = a.y; xv[r][2] = a.z; xv[r][3] = a.w; xv[r][4] = b.x; xv[r][5] = b.y; xv[r][6] = b.z; xv[r][7] = b.w;
                float s4 = 0.f;
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    if (normed) { s4 = fmaf(xv[r][i], xv[r][i], s4); mx = fmaxf(mx, fabsf(xv[r][i] * wv[r][i])); }
                    else mx = fmaxf(mx, fabsf(xv[r][i]));
                }
                ss += (double)s4;
            }
        }
        mx = warp_max(mx);
        if (normed) ss = warp_sum_d(ss);   // float64 across threads like the reference's float64 sum, go/quant.go:598-603
        if (lane == 0) { sh.mx_red[warp] = mx; sh.ss_red[warp] = ss; }
        tl_bar<TL_CONSUMERS>();
        float inv = 1.f;
        {
            float m2 = 0.f;
            double s2 = 0.0;
#pragma unroll
            for (int w = 0; w < TL_CW; w++) { m2 = fmaxf(m2, sh.mx_red[w]); s2 += sh.ss_red[w]; }
            mx = m2;
            if (normed) { inv = (float)(1.0 / sqrt(s2 / (double)cols + (double)A.eps)); mx *= inv; }
        }
        // power-of-two scale S: max|x| * S in [2^10, 2^11), so 16 * x * S stays inside fp16 and the lo terms keep 10+ bits
        int es = 264 - (int)((__float_as_uint(mx) >> 23) & 0xFFu);
        es = es < 27 ? 27 : (es > 227 ? 227 : es);
        const float S = __uint_as_float((uint32_t)es << 23);
        if (tid == 0) sh.post_scale[p & 1] = __uint_as_float((uint32_t)(274 - es) << 23);   // 2^20 / S
#pragma unroll
        for (int r = 0; r < TL_MAX_ITEMS; r++) {
            const int q = tid + r * TL_CONSUMERS;          // item q = elements [8q, 8q+8); whole warps agree on q < nitem_pad + 31
            const bool store = q < nitem_pad;              // zero fragments for the padding blocks of the last block group
            const int b = q >> 2, o = (q & 3) * 8;         // block, offset of my 8 elements inside it
            const int pos = o >> 4, ib = ((o & 15) >> 3) * 2;  // low | high nibble half, first of my two word indices
            float bs = 0.f;
            uint8_t *fb = xfrag + (size_t)(b >> 2) * 512 + (size_t)(2 * (b & 3)) * 64;
#pragma unroll
            for (int k = 0; k < 2; k++) {
                float v[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    float xx = 0.f;
                    if (q < nitem) {
                        xx = xv[r][4 * k + j];
                        if (normed) xx = xx * inv * wv[r][4 * k + j];   // x * inv * w, go/quant.go:604-606
                    }
                    v[j] = xx * S;
                    bs += v[j];
                    if (pos == 0) v[j] *= 16.f;
                }
                const uint32_t h0 = pack_h2(v[0], v[2]), h1 = pack_h2(v[1], v[3]);
                const float2 f0 = __half22float2(*reinterpret_cast<const __half2 *>(&h0)), f1 = __half22float2(*reinterpret_cast<const __half2 *>(&h1));
                const uint32_t l0 = pack_h2(v[0] - f0.x, v[2] - f0.y), l1 = pack_h2(v[1] - f1.x, v[3] - f1.y);
                const int reg = 4 * (ib + k) + 2 * pos;
                if (store) {
                    *reinterpret_cast<uint2 *>(fb + reg * 4) = make_uint2(h0, h1);         // hi column of this block
                    *reinterpret_cast<uint2 *>(fb + 64 + reg * 4) = make_uint2(l0, l1);    // lo column
                }
            }
            bs += __shfl_xor_sync(0xffffffffu, bs, 1);
            bs += __shfl_xor_sync(0xffffffffu, bs, 2);
            if (store && (q & 3) == 0) corr[b] = -7.62939453125e-6f * bs;   // -8 * 2^-20 * sum(x * S) over the block
        }
        tl_bar<TL_CONSUMERS>();
        if (tid == 0) TL_TRACE(p, 2);

        // ---- stream the band ----
        for (int c0 = 0; c0 < band; c0 += TL_TS, it++) {
            const int slot = it % TL_SLOTS;
            const int n = min(TL_TS, band - c0);
            mbar_wait(&sh.full_bar[slot], (it / TL_SLOTS) & 1);
            const uint8_t *sb = ring + (size_t)slot * TL_SLOT_BYTES;
            const int tl = 2 * warp;
            if (tl < n) {
                const int r0 = c0 + tl;
                const int q0 = rg_of(r0, nbg, magic);
                int B = r0 - q0 * nbg;
                float acc0 = 0.f, acc1 = 0.f;
                tile_dot(sb + (size_t)tl * TL_TILE, xfrag + (size_t)B * 512 + g * 64, corr[4 * B + t], xact, lane, xb, acc0, acc1);
                int e = 0;
                if (tl + 1 < n) {
                    B++;
                    if (B == nbg) {   // my second tile starts the next row group: flush the first
                        acc0 += __shfl_xor_sync(0xffffffffu, acc0, 1); acc0 += __shfl_xor_sync(0xffffffffu, acc0, 2);
                        acc1 += __shfl_xor_sync(0xffffffffu, acc1, 1); acc1 += __shfl_xor_sync(0xffffffffu, acc1, 2);
                        if (t == 0) { sh.red[slot][warp][0][g] = acc0; sh.red[slot][warp][0][g + 8] = acc1; }
                        acc0 = 0.f; acc1 = 0.f; B = 0; e = 1;
                    }
                    tile_dot(sb + (size_t)(tl + 1) * TL_TILE, xfrag + (size_t)B * 512 + g * 64, corr[4 * B + t], xact, lane, xb, acc0, acc1);
                }
                acc0 += __shfl_xor_sync(0xffffffffu, acc0, 1); acc0 += __shfl_xor_sync(0xffffffffu, acc0, 2);
                acc1 += __shfl_xor_sync(0xffffffffu, acc1, 1); acc1 += __shfl_xor_sync(0xffffffffu, acc1, 2);
                if (t == 0) { sh.red[slot][warp][e][g] = acc0; sh.red[slot][warp][e][g + 8] = acc1; }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sh.empty_bar[slot]);
        }
        if (tid == 0) TL_TRACE(p, 3);
    }
}

int launch_tiled(const TileArgs &a, int grid, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(decode_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TL_DYN_SMEM) != cudaSuccess) return -2;
        configured = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(TL_THREADS); cfg.dynamicSmemBytes = TL_DYN_SMEM; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative;  // all CTAs must be co-resident: they spin on each other
    at[0].val.cooperative = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, decode_tiled_kernel, a) == cudaSuccess ? 0 : -2;
}


int launch_tile_q4_0(const uint8_t *qs, const __half *d, int rows, int nb, uint8_t *tiles, int nbg, int rg_off, int rg_stride, cudaStream_t st) {
    const int n_rg_src = (rows + 15) / 16;
    const long long threads = (long long)n_rg_src * nbg * 32;
    tile_q4_0_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(reinterpret_cast<const uint4 *>(qs), d, rows, nb, tiles, nbg, rg_off, rg_stride, n_rg_src);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

}  // namespace nl
