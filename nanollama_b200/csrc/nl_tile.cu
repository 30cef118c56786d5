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

__global__ void tile_q8_0_kernel(const uint4 *__restrict__ qs, const __half *__restrict__ d, int rows, int nb, uint8_t *__restrict__ tiles,
                                 int nbg, int rg_off, int rg_stride, int n_rg_src) {
    constexpr int TILE = TileCfg<NL_Q8_0>::TILE;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = (int)(gid & 31);
    const long long tile = gid >> 5;
    if (tile >= (long long)n_rg_src * nbg) return;
    const int R = (int)(tile / nbg), B = (int)(tile % nbg);
    const int g = lane >> 2, t = lane & 3;
    const int blk = 4 * B + t;
    uint8_t *dst = tiles + ((size_t)(rg_off + R * rg_stride) * nbg + B) * TILE;
    const uint4 z = make_uint4(0, 0, 0, 0);
    uint4 a0 = z, a1 = z, b0 = z, b1 = z;
    unsigned short d0 = 0, d1 = 0;
    const int r0 = 16 * R + g, r1 = r0 + 8;
    if (blk < nb) {
        if (r0 < rows) { a0 = qs[((size_t)r0 * nb + blk) * 2]; a1 = qs[((size_t)r0 * nb + blk) * 2 + 1]; d0 = __half_as_ushort(d[(size_t)r0 * nb + blk]); }
        if (r1 < rows) { b0 = qs[((size_t)r1 * nb + blk) * 2]; b1 = qs[((size_t)r1 * nb + blk) * 2 + 1]; d1 = __half_as_ushort(d[(size_t)r1 * nb + blk]); }
    }
    reinterpret_cast<uint4 *>(dst)[lane] = a0;
    reinterpret_cast<uint4 *>(dst + 512)[lane] = a1;
    reinterpret_cast<uint4 *>(dst + 1024)[lane] = b0;
    reinterpret_cast<uint4 *>(dst + 1536)[lane] = b1;
    reinterpret_cast<uint32_t *>(dst + 2048)[lane] = (uint32_t)d0 | ((uint32_t)d1 << 16);
}

// ---- device helpers ----
// build-time variants of the streaming loop's inner product (tools/build_variants.py A/Bs them):
//   NL_TL_ACC     accumulator chains per tile: 2 = one per nibble half (4 dependent MMAs each), 4 / 8 = shorter chains + more adds
//   NL_TL_MMA_VOL 1 = the MMAs are volatile asm (kept in program order), 0 = the compiler may interleave them with the loads around them
#ifndef NL_TL_ACC
#define NL_TL_ACC 2
#endif
#ifndef NL_TL_MMA_VOL
#define NL_TL_MMA_VOL 1
#endif
#ifndef NL_TL_DBG_SKIPMATH
#define NL_TL_DBG_SKIPMATH 0
#endif
#ifndef NL_TL_FIN_SPIN
#define NL_TL_FIN_SPIN 0
#endif
#if NL_TL_MMA_VOL
#define TL_MMA_ASM asm volatile
#else
#define TL_MMA_ASM asm
#endif
__device__ __forceinline__ void mma_f16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    TL_MMA_ASM("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
    const __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<const uint32_t *>(&h);
}
template <int NT> __device__ __forceinline__ void tl_bar() { asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory"); }
// CTA b's share [u0, u1) of `units` distribution units: floor(units * b / G) without a divide (three roles evaluate this per phase, the
// math warps on the critical path; a 64-bit divide is a subroutine of hundreds of cycles).  gmagic = ceil(2^32 / G), exact while
// units * (b + 1) < 2^32 / G (launch_tiled checks the bound); G == 1 keeps everything.
__device__ __forceinline__ void band_of(int units, int b, int G, unsigned int gmagic, int &u0, int &u1) {
    if (G == 1) { u0 = 0; u1 = units; return; }
    u0 = (int)__umulhi((unsigned)(units * b), gmagic);
    u1 = (int)__umulhi((unsigned)(units * (b + 1)), gmagic);
}
// tile index inside a band -> row group (magic = ceil(2^32 / nbg); exact for r < 2^32 / nbg)
__device__ __forceinline__ int rg_of(int r, int nbg, unsigned int magic) { return nbg == 1 ? r : (int)__umulhi((unsigned)r, magic); }
template <typename T> __device__ __forceinline__ T *ldg_ptr(T *const *p) {
    return reinterpret_cast<T *>(__ldg(reinterpret_cast<const unsigned long long *>(p)));
}
#define TL_TRACE(p, k) do { if ((MODE == 0 || NL_TL_FINE_TRACE) && A.trace) A.trace[((size_t)blockIdx.x * A.n_phases + (p)) * 8 + (k)] = gtime(); } while (0)
// fine-grained forensics (compile with -DNL_TL_FINE_TRACE=1; off by default: the stamps cost registers in the phase loop): SM cycle
// counter (clock64) stamps, 16 per (CTA, phase); see tools/trace_fine.py for the slot meanings
#ifndef NL_TL_FINE_TRACE
#define NL_TL_FINE_TRACE 0
#endif
#if NL_TL_FINE_TRACE
#define TL_CK(p, k) do { if ((MODE == 0 || NL_TL_FINE_TRACE) && A.trace2) A.trace2[((size_t)blockIdx.x * A.n_phases + (p)) * 16 + (k)] = (unsigned long long)clock64(); } while (0)
#define CK_AT(k) do { if (ck) ck[k] = (unsigned long long)clock64(); } while (0)
#else
#define TL_CK(p, k) do { } while (0)
#define CK_AT(k) do { } while (0)
#endif
// build-time variants of the register allocation around the streaming loop (tools/build_variants.py A/Bs them)
#ifndef NL_TL_STREAM_INLINE
#define NL_TL_STREAM_INLINE 1
#endif
#if NL_TL_STREAM_INLINE
#define TL_STREAM_CALL __forceinline__
#else
#define TL_STREAM_CALL __noinline__
#endif
#ifndef NL_TL_ATTN_INLINE
#define NL_TL_ATTN_INLINE 0
#endif
#if NL_TL_ATTN_INLINE
#define TL_ATTN_CALL __forceinline__
#else
#define TL_ATTN_CALL __noinline__
#endif

// ---- flagged pairs: the un-normalised partials of a split attention item travel as 8-byte {value, flag} pairs written by ONE store, so a
// reader that sees the expected flag has the value too (no fence, no counter).  flag = epoch * (n_phases + 1) + producing phase + 1.
__device__ __forceinline__ void st_ll(void *base, int idx, float v, unsigned int flag) {
    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(reinterpret_cast<uint2 *>(base) + idx), "r"(__float_as_uint(v)), "r"(flag) : "memory");
}
__device__ __forceinline__ float ld_ll_wait(const void *base, int idx, unsigned int want) {
    unsigned int v, f;
    do {
        asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v), "=r"(f) : "l"(reinterpret_cast<const uint2 *>(base) + idx) : "memory");
    } while (f != want);
    return __uint_as_float(v);
}
// ---- polled activations (TileArgs::poll, single GPU): every activation vector of a token is written exactly ONCE (one vector per
// layer and producer in an arena that a memset node fills with TL_SENT in front of every launch), so an element that no longer reads
// TL_SENT IS the value: consumers poll the data itself.  No grid barrier, no fence, no flag bytes -- a phase boundary costs one
// store -> L2 -> load trip instead of store, fence, arrive, poll, load.  Every GEMV phase needs ALL outputs of the phase before it, so
// the data flow alone orders what the barriers ordered; single-use buffers leave no write-after-read hazard.  A computed value with
// the sentinel's bit pattern (one particular NaN) is stored as the canonical NaN instead.
constexpr unsigned int TL_SENT = 0xFFFFFFFFu;
__device__ __forceinline__ void st_poll(float *base, int idx, float v) {
    unsigned int b = __float_as_uint(v);
    if (b == TL_SENT) b = 0x7FFFFFFFu;
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(base + idx), "r"(b) : "memory");
}
__device__ __forceinline__ float ld_poll(const float *base, int idx) {
    unsigned int v;
    do {
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(base + idx) : "memory");
    } while (v == TL_SENT);
    return __uint_as_float(v);
}
__device__ __forceinline__ void ld_poll2(const float *base, int i0, int i1, float &v0, float &v1) {   // both looks in flight together
    unsigned int a, b;
    do {
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%2];\n\tld.relaxed.gpu.global.u32 %1, [%3];" : "=r"(a), "=r"(b) : "l"(base + i0), "l"(base + i1) : "memory");
    } while (a == TL_SENT || b == TL_SENT);
    v0 = __uint_as_float(a); v1 = __uint_as_float(b);
}
__device__ __forceinline__ uint4 ld_vol_v4(const void *p) {
    uint4 v;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint2 ld_vol_v2(const void *p) {
    uint2 v;
    asm volatile("ld.relaxed.gpu.global.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
    return v;
}
// 8 consecutive floats of a polled vector as raw words (two 16-byte loads)
__device__ __forceinline__ void ld_item(const float *base, int q, unsigned int (&w)[8]) {
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%8];\n\tld.relaxed.gpu.global.v4.u32 {%4,%5,%6,%7}, [%8+16];"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "l"(base + 8 * (size_t)q) : "memory");
}
__device__ __forceinline__ void ld_item_sys(const float *base, int q, unsigned int (&w)[8]) {   // (vectors a peer GPU writes over NVLink)
    asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%8];\n\tld.relaxed.sys.global.v4.u32 {%4,%5,%6,%7}, [%8+16];"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "l"(base + 8 * (size_t)q) : "memory");
}
__device__ __forceinline__ void st_poll_sys(float *p, float v) {   // a polled element in a PEER's arena
    unsigned int b = __float_as_uint(v);
    if (b == TL_SENT) b = 0x7FFFFFFFu;
    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(b) : "memory");
}
__device__ __forceinline__ bool item_has_sent(const unsigned int (&w)[8]) {
    return max(max(max(w[0], w[1]), max(w[2], w[3])), max(max(w[4], w[5]), max(w[6], w[7]))) == TL_SENT;   // the sentinel is the largest word
}
__device__ __forceinline__ bool has_sent(const uint4 &a, const uint4 &b) {
    return a.x == TL_SENT || a.y == TL_SENT || a.z == TL_SENT || a.w == TL_SENT || b.x == TL_SENT || b.y == TL_SENT || b.z == TL_SENT || b.w == TL_SENT;
}

// Grid barrier of phase p.  Single GPU: counters zeroed before every launch, target = grid.  Tensor parallel: counters live in the
// IPC window, are never reset (target = epoch x grid x ranks-that-arrive) and a phase whose outputs go to the peers (`cross`) is
// arrived at on EVERY rank's counter after a system-scope fence, so passing it means every rank's partials have landed here.
__device__ __forceinline__ void tl_arrive(const TileArgs &A, int p, bool cross) {
    if (A.tp <= 1) { phase_arrive(A.bar, p); return; }
    if (!cross) { phase_arrive(A.bar, p); return; }
    __threadfence_system();
    for (int r = 0; r < A.tp; r++)
        asm volatile("red.relaxed.sys.global.add.u32 [%0], 1;" ::"l"(reinterpret_cast<unsigned int *>(A.peers.win[r] + A.bar_off) + p) : "memory");
}
__device__ __forceinline__ void tl_wait(const TileArgs &A, int p, bool cross, unsigned int G, unsigned int epoch) {
    if (A.tp <= 1) { phase_wait(A.bar, p, G); return; }
    const unsigned int target = epoch * G * (cross ? (unsigned)A.tp : 1u);
    if (cross) { while ((int)(ld_acquire_sys(A.bar + p) - target) < 0) {} }
    else { while ((int)(ld_acquire(A.bar + p) - target) < 0) { __nanosleep(20); } }
}

// shared-memory accessors on raw 32-bit shared addresses (all address arithmetic stays in 32-bit integer registers, computed once)
__device__ __forceinline__ uint4 lds128(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts32f(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ void mbar_wait_u(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
// wait with a suspend-time hint: the hardware parks the warp until the phase completes (or the hint elapses) instead of letting it
// re-issue try_wait back to back.  The copy and finishing warps wait most of the time; spinning, they executed 7 % of all the kernel's
// instructions (ncu source counters) on the two SM sub-partitions they share with math warps.
__device__ __forceinline__ void mbar_wait_parked(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity), "r"(20000u) : "memory");
}
__device__ __forceinline__ void mbar_arrive_u(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }

// B fragments of one block group: only the lane that owns (block column, hi|lo) of B holds data, every other lane keeps zeros
__device__ __forceinline__ void load_xb(uint32_t xf_lane, int B, bool xact, uint32_t (&xb)[16]) {
    if (xact) {
#pragma unroll
        for (int i = 0; i < 4; i++) { const uint4 v = lds128(xf_lane + (uint32_t)B * (uint32_t)TL_XBG + 16u * i); xb[4 * i] = v.x; xb[4 * i + 1] = v.y; xb[4 * i + 2] = v.z; xb[4 * i + 3] = v.w; }
    }
}
// One tile: 16 rows x 4 blocks.  acc0 / acc1: this lane's running sums for (row g, block column t) and (row g+8, t).
// tile_lane = shared address of the tile + 16 * lane.  rec_addr: this block's record {corr_lo, pb_lo, corr_hi, pb_hi}: the low 16
// elements (low nibbles, accumulators c) and the high 16 (high nibbles, e) carry their own power-of-two operand scale, so that the
// producer of a 16-row group can publish its half block without knowing the other half (see "producer-side fragments").
__device__ __forceinline__ void tile_finish(const float (&c)[4], const float (&e)[4], uint32_t dd, uint32_t rec_addr, float &acc0, float &acc1) {
    float cl, pl, ch, ph;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(cl), "=f"(pl), "=f"(ch), "=f"(ph) : "r"(rec_addr));
    const float2 df = __half22float2(*reinterpret_cast<const __half2 *>(&dd));
    // per half: pb * ((hi + lo columns) - zero_point * sum(x) of the half)
    const float v0 = fmaf(ph, (e[0] + e[1]) + ch, pl * ((c[0] + c[1]) + cl));
    const float v1 = fmaf(ph, (e[2] + e[3]) + ch, pl * ((c[2] + c[3]) + cl));
    acc0 = fmaf(df.x, v0, acc0);
    acc1 = fmaf(df.y, v1, acc1);
}
__device__ __forceinline__ void tile_dot(uint32_t tile_lane, uint32_t d_lane, uint32_t rec_addr, const uint32_t (&xb)[16], float &acc0, float &acc1) {
    const uint4 wa4 = lds128(tile_lane);
    const uint4 wb4 = lds128(tile_lane + 512u);
    const uint32_t dd = lds32(d_lane);
    const uint32_t wa[4] = {wa4.x, wa4.y, wa4.z, wa4.w}, wb[4] = {wb4.x, wb4.y, wb4.z, wb4.w};
    constexpr int NCH = NL_TL_ACC / 2;   // chains per nibble half
    float cc[NCH][4], ee[NCH][4];
#pragma unroll
    for (int k = 0; k < NCH; k++) {
#pragma unroll
        for (int j = 0; j < 4; j++) { cc[k][j] = 0.f; ee[k][j] = 0.f; }
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const uint32_t a8 = wa[i] >> 8, b8 = wb[i] >> 8;
        // elements 4i..4i+3 (low nibbles): n * 2^-24 as fp16 subnormals; B carries 16 * x * S
        mma_f16(cc[i % NCH], wa[i] & 0x000F000Fu, wb[i] & 0x000F000Fu, a8 & 0x000F000Fu, b8 & 0x000F000Fu, xb[4 * i], xb[4 * i + 1]);
        // elements 16+4i..16+4i+3 (high nibbles): 16n * 2^-24; B carries x * S
        mma_f16(ee[i % NCH], wa[i] & 0x00F000F0u, wb[i] & 0x00F000F0u, a8 & 0x00F000F0u, b8 & 0x00F000F0u, xb[4 * i + 2], xb[4 * i + 3]);
    }
    float c[4], e[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        if constexpr (NCH == 1) { c[j] = cc[0][j]; e[j] = ee[0][j]; }
        else if constexpr (NCH == 2) { c[j] = cc[0][j] + cc[1][j]; e[j] = ee[0][j] + ee[1][j]; }
        else { c[j] = (cc[0][j] + cc[1][j]) + (cc[2][j] + cc[3][j]); e[j] = (ee[0][j] + ee[1][j]) + (ee[2][j] + ee[3][j]); }
    }
    tile_finish(c, e, dd, rec_addr, acc0, acc1);
}

// Q8_0 tile: eight MMAs, each on word i (elements 4i..4i+3) of the lane's block; B fragment registers 2i, 2i+1.  Words 0-3 (elements
// 0-15) go through c, words 4-7 (elements 16-31) through e: the same half-block split as Q4_0.
__device__ __forceinline__ void tile_dot_q8(uint32_t tile_lane, uint32_t d_lane, uint32_t rec_addr, const uint32_t (&xb)[16], float &acc0, float &acc1) {
    const uint4 a_lo = lds128(tile_lane), a_hi = lds128(tile_lane + 512u), b_lo = lds128(tile_lane + 1024u), b_hi = lds128(tile_lane + 1536u);
    const uint32_t dd = lds32(d_lane);
    const uint32_t wa[8] = {a_lo.x, a_lo.y, a_lo.z, a_lo.w, a_hi.x, a_hi.y, a_hi.z, a_hi.w};
    const uint32_t wb[8] = {b_lo.x, b_lo.y, b_lo.z, b_lo.w, b_hi.x, b_hi.y, b_hi.z, b_hi.w};
    float c[4] = {0.f, 0.f, 0.f, 0.f}, e[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint32_t ua = wa[i] ^ 0x80808080u, ub = wb[i] ^ 0x80808080u;   // int8 -> q + 128 in every byte
        if (i >= 4) mma_f16(e, ua & 0x00FF00FFu, ub & 0x00FF00FFu, (ua >> 8) & 0x00FF00FFu, (ub >> 8) & 0x00FF00FFu, xb[2 * i], xb[2 * i + 1]);
        else mma_f16(c, ua & 0x00FF00FFu, ub & 0x00FF00FFu, (ua >> 8) & 0x00FF00FFu, (ub >> 8) & 0x00FF00FFu, xb[2 * i], xb[2 * i + 1]);
    }
    tile_finish(c, e, dd, rec_addr, acc0, acc1);
}

// ---- producer-side fragments: one 16-lane group holds 16 consecutive elements of the NEXT GEMV's input (a half block: element
// e0 + lane, e0 a multiple of 16) and publishes them in that GEMV's final form (nl_tile.cuh, TL_IMG_BG).  y = the value the GEMV multiplies
// (x o w for a normed input), x2 = this lane's contribution to the RMSNorm sum of squares.  All 32 lanes call it; `sel` = 0: this
// half-warp stores the hi column, 1: the lo column (the two half-warps hold the same 16 values), lane16 = lane & 15.
// Layout facts (the consumer-side conversion input_frags_body writes the same bytes): block b = e0 >> 5 sits at (b >> 2) * TL_IMG_BG +
// (b & 3) * 128, hi column first (64 B), lo column at +64; Q4_0: element e' (0..15) of half `pos` is halfword (e' >> 1) & 1 of word
// 4 * (e' >> 2) + 2 * pos + (e' & 1); Q8_0: halfword (e' >> 1) & 1 of word 8 * pos + 2 * (e' >> 2) + (e' & 1).  A word therefore pairs
// lanes L and L + 2 with (L & 2) == 0.  Records: 16 bytes per half block behind the block group's 512 fragment bytes.
// sel = 2: this 16-lane group stores both columns (attention epilogue: the two half-warps hold different half blocks).
template <int TYPE>
__device__ __forceinline__ void publish_half_block(uint8_t *img, int e0, int lane16, int sel, float y, float x2) {
    float mx = fabsf(y);
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    // Q4_0: max|y| * S in [2^10, 2^11) (the low-nibble operands carry 16 * y * S); Q8_0: [2^14, 2^15)
    int es = (TYPE == NL_Q8_0 ? 268 : 264) - (int)((__float_as_uint(mx) >> 23) & 0xFFu);
    es = es < 27 ? 27 : (es > 227 ? 227 : es);
    const float S = __uint_as_float((uint32_t)es << 23);
    const int b = e0 >> 5, pos = (e0 >> 4) & 1;
    float v = y * S, bs = v, ss = x2;
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) { bs += __shfl_xor_sync(0xffffffffu, bs, o); ss += __shfl_xor_sync(0xffffffffu, ss, o); }
    if (TYPE == NL_Q4_0 && pos == 0) v *= 16.f;
    const __half hh = __float2half_rn(v);
    const __half hl = __float2half_rn(v - __half2float(hh));
    const uint32_t mine = (uint32_t)__half_as_ushort(hh) | ((uint32_t)__half_as_ushort(hl) << 16);   // (hi, lo) of my element
    const uint32_t other = __shfl_xor_sync(0xffffffffu, mine, 2);
    uint8_t *bg = img + (size_t)(b >> 2) * TL_IMG_BG;
    if ((lane16 & 2) == 0) {
        const int i = lane16 >> 2, j = lane16 & 1;
        const int word = (TYPE == NL_Q8_0) ? 8 * pos + 2 * i + j : 4 * i + 2 * pos + j;
        uint8_t *wp = bg + (b & 3) * 128 + word * 4;
        if (sel != 1) asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(wp), "r"((mine & 0xFFFFu) | (other << 16)) : "memory");
        if (sel != 0) asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(wp + 64), "r"((mine >> 16) | (other & 0xFFFF0000u)) : "memory");
    }
    if (lane16 == 1 && sel != 1) {   // the half block's record: one 16-byte store
        unsigned int cw = __float_as_uint(-7.62939453125e-6f * bs);   // -zero_point * 2^-k * sum(y * S): 8 * 2^-20 = 128 * 2^-24 = 2^-17
        if (cw == 0xFFFFFFFFu) cw = 0x7FFFFFFFu;
        const unsigned int pw = (uint32_t)((TYPE == NL_Q8_0 ? 278 : 274) - es) << 23;   // 2^k / S
        unsigned int sw = __float_as_uint(ss);
        if (sw == 0xFFFFFFFFu) sw = 0x7FFFFFFFu;
        asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(bg + 512 + ((b & 3) * 2 + pos) * 16), "r"(cw), "r"(pw), "r"(sw), "r"(0u) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Attention phase (go/model.go:530-587): item = (kv head, split of the positions).  The K/V rows of earlier positions do not depend on
// the current token, so an item's first 64 rows are fetched into shared memory (cp.async) BEFORE the grid barrier that publishes
// q / k / v; after the barrier one round trip brings q, k, v, everything else is on chip.  Short contexts use one split per
// 64 positions (a single split writes the final output directly, no cross-CTA combine); longer ones use up to at.nsplit splits
// whose un-normalised partials (flash-decoding) are folded by the split that arrives last (fixed order => deterministic).
constexpr int TA_CH = 96;   // positions per pass (one pass per item up to 9 x 96 positions of context)
struct AttnT {              // lives in the fragment buffer during the attention phase
    float q[MG_MAX_GROUP][64];
    float knew[64], vnew[64];
    float p[MG_MAX_GROUP][TA_CH];
    float Ks[TA_CH][68];    // 272-byte rows: 16-byte aligned, and 8 consecutive rows hit 8 different bank groups (reused for the
                            // per-thread PV partials once the last pass is over: TL_CONSUMERS * 4 floats)
    float Vs[TA_CH][64];
    float m_run[MG_MAX_GROUP], l_run[MG_MAX_GROUP], corr[MG_MAX_GROUP];
    int is_last;
};
static_assert(sizeof(AttnT) <= TL_XFRAG_BYTES, "attention scratch must fit the fragment buffer");
static_assert(TA_CH * 68 >= TL_CONSUMERS * 4 && TA_CH <= 96, "PV partials alias the K rows; the softmax step covers 3 x 32 positions");
static_assert(TL_CW == 16, "the finishing warp sums 8 partials per half-warp");

__device__ __forceinline__ void cp_async16(void *dst_smem, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// item = (hpi consecutive q heads of one kv head, split of the positions)
struct AttnItem { int kvh, h0, nh, wr, sp, nse, t_begin, t_end; };
struct TlShared {
    uint64_t full_bar[TL_SLOTS], empty_bar[TL_SLOTS], free_bar[TL_SLOTS];
    float red[TL_SLOTS][TL_CW][2][16];
    double ss_red[TL_CW];
    TilePhase ph[3];
    MegaAttn at;              // copy of TileArgs::at: out-of-line device functions cannot take the address of a kernel parameter for free
    unsigned long long *trace, *trace2;   // (same: TileArgs::trace / trace2, n_phases, poll_ns, tp, dim)
    const float *ar_mine;     // tensor parallel: this rank's exchange area
    int n_phases, poll_ns, tp, dim;
    // the attention phase's shape is the same in every layer of a token: position, splits, q heads per item, item count and this CTA's
    // first item are worked out once per launch (a global load and a handful of integer divides per layer otherwise)
    int a_pos, a_nse, a_hpi, a_items;
    AttnItem a_item0;
};

// item = (hpi consecutive q heads of one kv head, split of the positions).  hpi = the whole GQA group shares one pass over the K/V rows;
// smaller hpi (attn_hpi) spreads a short context over more CTAs -- the rows come from L2 anyway and an item's time is per-head work.
__device__ __forceinline__ int attn_hpi(const MegaAttn &at, int nse, int G, int forced) {
    const int group = at.n_heads / at.n_kv_heads;
    if (forced > 0) return forced >= group || group % forced ? group : forced;
    for (int hpi = 1; hpi < group; hpi++)
        if (group % hpi == 0 && (at.n_heads / hpi) * nse <= G) return hpi;
    return group;
}
__device__ __forceinline__ AttnItem attn_locate(const MegaAttn &at, int item, int n, int nse, int hpi) {
    AttnItem I;
    const int group = at.n_heads / at.n_kv_heads, vk = item / nse;
    I.h0 = vk * hpi; I.nh = hpi; I.kvh = I.h0 / group; I.wr = (I.h0 - I.kvh * group) == 0;   // one item per kv head writes the new K/V row
    I.sp = item - vk * nse; I.nse = nse;
    const int per = (n + nse - 1) / nse;
    I.t_begin = min(I.sp * per, n); I.t_end = min(I.t_begin + per, n);
    return I;
}
// rows [t0, t0 + cn) below `pos` of the K and V cache of one kv head -> S.Ks / S.Vs (asynchronous)
__device__ __forceinline__ void attn_fetch(const float *kc, const float *vc, int kvd, int kvh, int t0, int cn, int pos, AttnT &S, int tid) {
    for (int idx = tid; idx < cn * 32; idx += TL_CONSUMERS) {
        const int tl = idx >> 5, f = idx & 31, t = t0 + tl;
        if (t < pos) {
            if (f < 16) cp_async16(&S.Ks[tl][4 * f], kc + (size_t)t * kvd + kvh * 64 + 4 * f);
            else cp_async16(&S.Vs[tl][4 * (f - 16)], vc + (size_t)t * kvd + kvh * 64 + 4 * (f - 16));
        }
    }
}

// (few scalar arguments: they travel in registers; the rest -- this layer's q | k | v vector (polled element by element when `poll`), its
// attention output, the fragment image of the o-projection's input -- is read from the phase descriptor in shared memory)
template <int TYPE, int MODE>
__device__ TL_ATTN_CALL void attn_item_tiled(const TlShared &sh, AttnT &S, int item, int nse, int hpi, int pos, int p,
                                             bool prefetched, bool poll, unsigned int oflag) {
    const MegaAttn &at = sh.at;
    const TilePhase &P = sh.ph[p % 3];
    const float *qkv = P.x;
    float *ao = P.out;
    const int layer = P.layer;
    const int tid = threadIdx.x;
    const AttnItem I = item == (int)blockIdx.x ? sh.a_item0 : attn_locate(at, item, pos + 1, nse, hpi);
    unsigned long long *trace = ((MODE == 0 || NL_TL_FINE_TRACE) && sh.trace && item == (int)blockIdx.x) ? sh.trace + ((size_t)blockIdx.x * sh.n_phases + p) * 8 : nullptr;
    [[maybe_unused]] unsigned long long *ck = (NL_TL_FINE_TRACE && sh.trace2 && item == (int)blockIdx.x && tid == 0) ? sh.trace2 + ((size_t)blockIdx.x * sh.n_phases + p) * 16 : nullptr;
    constexpr int HD = 64, HALF = 32;
    const int group = I.nh, kvd = at.n_kv_heads * HD;   // "group": the q heads of THIS item (I.h0 .. I.h0 + group - 1)
    const int warp = tid >> 5, lane = tid & 31;
    const int kvh = I.kvh;
    float *kc = at.kcache + (size_t)layer * at.seq_len * kvd, *vc = at.vcache + (size_t)layer * at.seq_len * kvd;
    const bool owner = pos >= I.t_begin && pos < I.t_end;   // exactly one split per kv head holds the new position

    // ---- RoPE of the group's q heads and of the new k (go/model.go:449-477, :530-539); v as it is
    if (tid < (group + 1) * HALF) {
        const int hh = tid >> 5, i = tid & 31;
        const bool isk = hh == group;
        // q | k | v are one vector [H*hd + 2*kvd] (at.q = base; at.k / at.v are element offsets from it)
        const int e0 = isk ? (int)(at.k - at.q) + kvh * HD : (I.h0 + hh) * HD;
        float x0, x1;
        if (poll) ld_poll2(qkv, e0 + i, e0 + i + HALF, x0, x1);
        else { x0 = __ldcg(qkv + e0 + i); x1 = __ldcg(qkv + e0 + i + HALF); }
        const float c = __ldg(at.cos_t + (size_t)pos * HALF + i), sn = __ldg(at.sin_t + (size_t)pos * HALF + i);
        float r0, r1;
        if (!at.conj) { r0 = x0 * c - x1 * sn; r1 = x0 * sn + x1 * c; }
        else { r0 = x0 * c + x1 * sn; r1 = -x0 * sn + x1 * c; }
        float *dst = isk ? S.knew : S.q[hh];
        dst[i] = r0; dst[i + HALF] = r1;
    } else if (tid >= TL_CONSUMERS - HD) {
        const int ev = (int)(at.v - at.q) + kvh * HD + tid - (TL_CONSUMERS - HD);
        S.vnew[tid - (TL_CONSUMERS - HD)] = poll ? ld_poll(qkv, ev) : __ldcg(qkv + ev);
    }
    if (tid < group) { S.m_run[tid] = -INFINITY; S.l_run[tid] = 0.f; S.corr[tid] = 0.f; }
    CK_AT(2);
    tl_bar<TL_CONSUMERS>();
    if (tid == 0 && trace) trace[5] = gtime();   // q / k / v in shared memory
    if (at.qk_norm) {
        if (warp <= group) {  // RMSNormBare, go/quant.go:584-594
            float *vec = warp == group ? S.knew : S.q[warp];
            double ss = 0.0;
            for (int i = lane; i < HD; i += 32) ss += (double)vec[i] * (double)vec[i];
            ss = warp_sum_d(ss);
            const float inv = (float)(1.0 / sqrt(ss / (double)HD + (double)at.eps));
            for (int i = lane; i < HD; i += 32) vec[i] *= inv;
        }
        tl_bar<TL_CONSUMERS>();
    }
    CK_AT(3);
    if (owner && I.wr && tid < HD) {   // KV write, go/model.go:552-554
        kc[(size_t)pos * kvd + kvh * HD + tid] = S.knew[tid];
        vc[(size_t)pos * kvd + kvh * HD + tid] = S.vnew[tid];
    }

    const int gthreads = group * HD;                       // epilogue: one thread per (head of the group, dim)
    const int pthreads = group * 16;                       // PV: one thread per (head, 4 dims), positions interleaved over nparts thread sets
    const int nparts = TL_CONSUMERS / pthreads;
    const int part = tid / pthreads, prem = tid - part * pthreads, hh = prem >> 4, quad = prem & 15;
    float4 acc4 = make_float4(0.f, 0.f, 0.f, 0.f);
    int pass = 0;
    for (int c0 = I.t_begin; c0 < I.t_end; c0 += TA_CH, pass++) {
        const int cn = min(TA_CH, I.t_end - c0);
        if (!(prefetched && pass == 0)) attn_fetch(kc, vc, kvd, kvh, c0, cn, pos, S, tid);
        if (owner && pos >= c0 && pos < c0 + cn) {          // the new position's row comes from this token's k / v
            if (tid < HD) S.Ks[pos - c0][tid] = S.knew[tid];
            else if (tid < 2 * HD) S.Vs[pos - c0][tid - HD] = S.vnew[tid - HD];
        }
        cp_async_wait_all();
        tl_bar<TL_CONSUMERS>();
        if (pass == 0) CK_AT(4);
        {   // scores: one lane per position, one warp per (head, 32 positions) -- the whole 64-element dot product in the lane, no
            // shuffles (the SM retires about one warp shuffle per clock: sixty per warp on sixteen warps were ~1000 cycles of this phase).
            // K rows are 272 bytes apart: eight consecutive rows cover the 32 banks once, the 16-byte loads of a quarter warp do not
            // conflict; the q loads are broadcasts.
            constexpr int NCH = (TA_CH + 31) / 32;
            for (int hw = warp; hw < NCH * group; hw += TL_CW) {
                const int h2 = hw / NCH, tl = (hw - h2 * NCH) * 32 + lane;
                if (tl < cn) {
                    float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
                    for (int i = 0; i < HD / 4; i++) {
                        const float4 kv = *reinterpret_cast<const float4 *>(&S.Ks[tl][4 * i]);
                        const float4 qv = *reinterpret_cast<const float4 *>(&S.q[h2][4 * i]);
                        d0 = fmaf(qv.x, kv.x, d0); d1 = fmaf(qv.y, kv.y, d1); d2 = fmaf(qv.z, kv.z, d2); d3 = fmaf(qv.w, kv.w, d3);
                    }
                    S.p[h2][tl] = ((d0 + d1) + (d2 + d3)) * at.scale;
                }
            }
        }
        tl_bar<TL_CONSUMERS>();
        if (pass == 0) CK_AT(5);
        if (warp < group) {  // running softmax statistics of head `warp` (flash-decoding form of go/quant.go:610-626)
            const float s0 = lane < cn ? S.p[warp][lane] : -INFINITY, s1 = lane + 32 < cn ? S.p[warp][lane + 32] : -INFINITY;
            const float s2 = lane + 64 < cn ? S.p[warp][lane + 64] : -INFINITY;
            const float mx = warp_max(fmaxf(fmaxf(s0, s1), s2));
            const float m_old = S.m_run[warp], m_new = fmaxf(m_old, mx);
            const float e0 = lane < cn ? expf(s0 - m_new) : 0.f, e1 = lane + 32 < cn ? expf(s1 - m_new) : 0.f, e2 = lane + 64 < cn ? expf(s2 - m_new) : 0.f;
            if (lane < cn) S.p[warp][lane] = e0;
            if (lane + 32 < cn) S.p[warp][lane + 32] = e1;
            if (lane + 64 < cn) S.p[warp][lane + 64] = e2;
            const float sum = warp_sum((e0 + e1) + e2);
            if (lane == 0) {
                const float corr = (m_old == -INFINITY) ? 0.f : expf(m_old - m_new);
                S.corr[warp] = corr;
                S.l_run[warp] = S.l_run[warp] * corr + sum;
                S.m_run[warp] = m_new;
            }
        }
        tl_bar<TL_CONSUMERS>();
        if (pass == 0) CK_AT(6);
        if (part < nparts) {   // PV: thread = (part of the positions, head, 4 dims)
            const float cr = S.corr[hh];
            float4 a = make_float4(acc4.x * cr, acc4.y * cr, acc4.z * cr, acc4.w * cr);
            for (int tl = part; tl < cn; tl += nparts) {
                const float pw = S.p[hh][tl];
                const float4 v = *reinterpret_cast<const float4 *>(&S.Vs[tl][4 * quad]);
                a.x = fmaf(pw, v.x, a.x); a.y = fmaf(pw, v.y, a.y); a.z = fmaf(pw, v.z, a.z); a.w = fmaf(pw, v.w, a.w);
            }
            acc4 = a;
        }
        tl_bar<TL_CONSUMERS>();
        if (pass == 0) CK_AT(7);
        if (tid == 0 && trace && pass == 0) trace[2] = gtime();   // first pass done
    }
    float *pvs = &S.Ks[0][0];   // every pass is over (the loop ends on a barrier): the K rows become the PV partials
    *reinterpret_cast<float4 *>(&pvs[tid * 4]) = acc4;
    tl_bar<TL_CONSUMERS>();
    CK_AT(8);
    if (tid == 0 && trace) trace[6] = gtime();   // own positions done
    if (tid < gthreads) {
        const int hh = tid >> 6, dd = tid & 63;            // (shadows the PV mapping)
        float o = 0.f;
        for (int pp = 0; pp < nparts; pp++) o += pvs[(pp * pthreads + hh * 16 + (dd >> 2)) * 4 + (dd & 3)];
        const int h = I.h0 + hh;
        float M = S.m_run[hh], den = S.l_run[hh];
        if (I.nse > 1) {
            // un-normalised partials travel as flagged {value, flag} pairs: split 0 folds the others as they land (fixed order =>
            // deterministic), no fence, no counter
            const unsigned int pflag = oflag;
            if (I.sp != 0) {
                st_ll(at.part_acc, (h * at.nsplit + I.sp) * HD + dd, o, pflag);
                if (dd < 2) st_ll(at.part_ml, (h * at.nsplit + I.sp) * 2 + dd, dd == 0 ? M : den, pflag);
            } else {
                // split 0 folds the others in split order (deterministic), four at a time: the looks at one group's {max, sum} pairs and
                // partial outputs are all in flight together (one L2 round trip per group when the partials have landed), then merged
                // with the running-softmax rule
                const uint4 *mlp = reinterpret_cast<const uint4 *>(at.part_ml) + (size_t)h * at.nsplit;          // {M, flag, sum, flag} per split
                const uint2 *pap = reinterpret_cast<const uint2 *>(at.part_acc) + (size_t)h * at.nsplit * HD + dd;   // + s * HD
                for (int s0 = 1; s0 < I.nse; s0 += 4) {
                    uint4 ml[4];
                    uint2 pa[4];
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        if (s0 + j < I.nse) { ml[j] = ld_vol_v4(mlp + s0 + j); pa[j] = ld_vol_v2(pap + (size_t)(s0 + j) * HD); }
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        if (s0 + j >= I.nse) break;
                        while (ml[j].y != pflag || ml[j].w != pflag) ml[j] = ld_vol_v4(mlp + s0 + j);
                        while (pa[j].y != pflag) pa[j] = ld_vol_v2(pap + (size_t)(s0 + j) * HD);
                        const float ms = __uint_as_float(ml[j].x), ls = __uint_as_float(ml[j].z);
                        if (ls > 0.f) {
                            const float mn = fmaxf(M, ms);
                            const float wa = den > 0.f ? expf(M - mn) : 0.f, wb = expf(ms - mn);
                            den = fmaf(wb, ls, wa * den);
                            o = fmaf(wb, __uint_as_float(pa[j].x), wa * o);
                            M = mn;
                        }
                    }
                }
            }
        }
        if (I.nse == 1 || I.sp == 0) {   // (uniform per item; tid < gthreads is whole warps)
            const float r = o * (1.0f / den);
            if (ao) { if (poll) st_poll(ao, h * HD + dd, r); else ao[h * HD + dd] = r; }
            // the o-projection's input in its final form: every 16-lane group holds one half block of the attention output
            if (MODE != 1 && P.out_img) publish_half_block<TYPE>(P.out_img, h * HD + (dd & ~15), dd & 15, 2, r, 0.f);
        }
    }
    CK_AT(9);
    tl_bar<TL_CONSUMERS>();  // S is reused by the next item of this CTA
    CK_AT(10);
    if (tid == 0 && trace) trace[7] = gtime();   // outputs / partials stored
}


// Math-warp side of one GEMV phase: consume this CTA's band slot by slot.  Inlined into gemv_phase (out of line there), so that its registers (the B
// fragments, the shared-memory addresses) are allocated for the loop alone, not on top of the phase prologue's.
// My tiles of slot k are band tiles TS * k + TPW * warp (+1 when TPW == 2); their block group advances by TS mod nbg per slot.
// EVEN (Q4_0, nbg even -- every shape but the 576 / 640-column matrices of nano / micro): a warp's two tiles always belong to the same
// row group (bands start on row groups and slots on even tiles), so the loop body is two tiles in a straight line, one fold, one pair
// of stores; the general body flushes between the tiles when the second one starts the next row group.
template <int TYPE, bool EVEN>
__device__ TL_STREAM_CALL int stream_band(int band, int nbg, unsigned int magic, int it, uint32_t ring_u, uint32_t sh_u) {
    // (six arguments: they travel in registers; more of them went through the stack, i.e. through local memory in the hot loop)
    constexpr int TILE = TileCfg<TYPE>::TILE, TPW = TileCfg<TYPE>::TPW, TS = TPW * TL_CW, D_OFF = TileCfg<TYPE>::D_OFF;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t xfrag_u = ring_u + (uint32_t)TL_SLOTS * TL_SLOT_BYTES, corr_u = xfrag_u + (uint32_t)TL_XFRAG_BYTES;
    const uint32_t full_u = sh_u + (uint32_t)offsetof(TlShared, full_bar), empty_u = sh_u + (uint32_t)offsetof(TlShared, empty_bar);
    uint32_t red_lane = sh_u + (uint32_t)offsetof(TlShared, red) + (uint32_t)(warp * 2 * 16 + (lane >> 2)) * 4u;   // &red[0][warp][0][lane >> 2]
    const int g = lane >> 2, t = lane & 3;
    const bool xact = (t == (g >> 1));     // lane that holds B column g (block column g>>1, hi|lo = g&1)
    uint32_t xb[16];                       // B fragments of the tile at hand (ONE set, reloaded per tile: a second set spilled)
#pragma unroll
    for (int i = 0; i < 16; i++) xb[i] = 0u;
    const int t0 = TPW * warp;             // my first tile of a slot (the second one, t0 + 1, only when TPW == 2)
    uint32_t tile_lane0 = ring_u + (uint32_t)t0 * TILE + (uint32_t)lane * 16u;   // + slot * TL_SLOT_BYTES
    uint32_t d_lane0 = ring_u + (uint32_t)t0 * TILE + (uint32_t)D_OFF + (uint32_t)lane * 4u;
    uint32_t xf_lane = xfrag_u + (uint32_t)g * (uint32_t)TL_XCOL, rec_lane = corr_u + (uint32_t)t * 16u;
    asm volatile("" : "+r"(tile_lane0), "+r"(d_lane0), "+r"(xf_lane), "+r"(rec_lane), "+r"(red_lane));   // keep them in registers: no re-derivation per slot
    int B = t0 - rg_of(t0, nbg, magic) * nbg;
    const int stepB = TS - rg_of(TS, nbg, magic) * nbg;
    int cb = -1;                           // block group whose fragments xb holds
    for (int c0 = 0; c0 < band; c0 += TS, it++) {
        const uint32_t slot = (uint32_t)it % TL_SLOTS;
        const int n = min(TS, band - c0);
        mbar_wait_u(full_u + slot * 8u, ((uint32_t)it / TL_SLOTS) & 1u);
#if NL_TL_DBG_SKIPMATH   // (forensics, with NL_TILE_DBG=1: what the slot hand-over costs without the tile products)
        if (false) {
#else
        if (t0 < n) {
#endif
            const uint32_t so = slot * (uint32_t)TL_SLOT_BYTES, ro = red_lane + slot * (uint32_t)(TL_CW * 2 * 16 * 4);
            float acc0 = 0.f, acc1 = 0.f;
            if (B != cb) { load_xb(xf_lane, B, xact, xb); cb = B; }
            if constexpr (TYPE == NL_Q8_0) tile_dot_q8(tile_lane0 + so, d_lane0 + so, rec_lane + (uint32_t)B * 64u, xb, acc0, acc1);
            else tile_dot(tile_lane0 + so, d_lane0 + so, rec_lane + (uint32_t)B * 64u, xb, acc0, acc1);
            uint32_t eo = 0;
            if constexpr (TPW == 2 && EVEN) {   // (t0 + 1 < n and B + 1 < nbg always hold here)
                load_xb(xf_lane, B + 1, xact, xb); cb = B + 1;
                tile_dot(tile_lane0 + so + TILE, d_lane0 + so + TILE, rec_lane + (uint32_t)(B + 1) * 64u, xb, acc0, acc1);
            } else if (TPW == 2 && t0 + 1 < n) {
                int B1 = B + 1;
                if (B1 == nbg) {   // my second tile starts the next row group: flush the first
                    acc0 += __shfl_xor_sync(0xffffffffu, acc0, 1); acc0 += __shfl_xor_sync(0xffffffffu, acc0, 2);
                    acc1 += __shfl_xor_sync(0xffffffffu, acc1, 1); acc1 += __shfl_xor_sync(0xffffffffu, acc1, 2);
                    if (t == 0) { sts32f(ro, acc0); sts32f(ro + 32u, acc1); }
                    acc0 = 0.f; acc1 = 0.f; B1 = 0; eo = 64u;
                }
                if (B1 != cb) { load_xb(xf_lane, B1, xact, xb); cb = B1; }
                tile_dot(tile_lane0 + so + TILE, d_lane0 + so + TILE, rec_lane + (uint32_t)B1 * 64u, xb, acc0, acc1);
            }
            acc0 += __shfl_xor_sync(0xffffffffu, acc0, 1); acc1 += __shfl_xor_sync(0xffffffffu, acc1, 1);
            acc0 += __shfl_xor_sync(0xffffffffu, acc0, 2); acc1 += __shfl_xor_sync(0xffffffffu, acc1, 2);
            if (t == 0) { sts32f(ro + eo, acc0); sts32f(ro + eo + 32u, acc1); }
        }
        B += stepB;
        if (B >= nbg) B -= nbg;
        __syncwarp();
        if (lane == 0) mbar_arrive_u(empty_u + slot * 8u);
    }
    return it;
}

// Phase input -> fp16 hi/lo B fragments in shared memory (xfrag) + per-block corrections (corr); returns this thread's part of the sum
// of squares (RMSNorm).  mode 0: plain vector, 1: polled vector (looked at until no element reads as the sentinel), 2: tensor-parallel
// exchange (xprev + the ranks' partials in rank order).  ckrow: optional clock64 stamps of tid 0 (slot 15: globaltimer of "first item valid").
template <int TYPE, int mode>
__device__ __forceinline__ double input_frags_body(const float *px, const float *pnw, int nitem, int nitem_pad, uint8_t *xfrag, float2 *corr, int tid, int poll_ns,
                                                   const float *xprev, float *xnext, const float *xparts, int tp, int dim, bool xstore, unsigned long long *ckrow,
                                                   bool xparts_prev_poll = false) {
    // One pass, no grid-wide reduction in front of the conversion: every 32-element block gets its own power-of-two scale S_b
    // (max|y| * S_b in [2^10, 2^11), so 16 * y * S_b stays inside fp16 and the lo terms keep 10+ bits), applied back per block in
    // tile_dot; the RMSNorm scale (one scalar per vector) is applied by the finishing warp to the finished sums:
    // W . (inv * (x o w)) = inv * (W . (x o w)).  The float64 sum of squares is therefore off the critical path.
    double ss = 0.0;
    constexpr bool in_poll = mode == 1, in_exch = mode == 2, in_exch_poll = mode == 3;
    const bool normed = pnw != nullptr;
    float wv0[8];   // norm weights of my first item (items beyond the first only exist for dim > 4096: fetched where they are used)
#pragma unroll
    for (int i = 0; i < 8; i++) wv0[i] = 1.f;   // (defined on every path: left undefined, the compiler keeps the array in local memory)
    if (normed && tid < nitem) {   // static data: requested before the first look at x
        const float4 a = __ldg(reinterpret_cast<const float4 *>(pnw) + 2 * tid), b = __ldg(reinterpret_cast<const float4 *>(pnw) + 2 * tid + 1);
        wv0[0] = a.x; wv0[1] = a.y; wv0[2] = a.z; wv0[3] = a.w; wv0[4] = b.x; wv0[5] = b.y; wv0[6] = b.z; wv0[7] = b.w;
    }
    // polled input: one item at a time, looked at until it is complete (items held across the loop were spilled to local memory, which is
    // an L2 trip each with 12 KB of L1 left next to the ring; only 11008-column inputs have more than one item per thread anyway)
#pragma unroll
    for (int r = 0; r < TL_MAX_ITEMS; r++) {
        const int q = tid + r * TL_CONSUMERS;          // item q = elements [8q, 8q+8); whole warps agree on q < nitem_pad + 31
        if (r > 0 && (q & ~31) >= nitem_pad) break;     // (warp-uniform) nothing left for this warp: 4096-column inputs are one round
        const bool store = q < nitem_pad;              // zero fragments for the padding blocks of the last block group
        float y[8];
#pragma unroll
        for (int i = 0; i < 8; i++) y[i] = 0.f;
        if (q < nitem) {
            if (in_exch) {   // residual + every rank's partial of the row-split product, in rank order (the all-reduce)
                const float4 a = __ldcg(reinterpret_cast<const float4 *>(xprev) + 2 * q), b = __ldcg(reinterpret_cast<const float4 *>(xprev) + 2 * q + 1);
                y[0] = a.x; y[1] = a.y; y[2] = a.z; y[3] = a.w; y[4] = b.x; y[5] = b.y; y[6] = b.z; y[7] = b.w;
                const float *parts = xparts;
                for (int rr = 0; rr < tp; rr++) {
                    const float4 c = __ldcg(reinterpret_cast<const float4 *>(parts + (size_t)rr * dim) + 2 * q), d = __ldcg(reinterpret_cast<const float4 *>(parts + (size_t)rr * dim) + 2 * q + 1);
                    y[0] += c.x; y[1] += c.y; y[2] += c.z; y[3] += c.w; y[4] += d.x; y[5] += d.y; y[6] += d.z; y[7] += d.w;
                }
                if (xstore) {   // the new residual, read back at the next exchange
                    reinterpret_cast<float4 *>(xnext)[2 * q] = make_float4(y[0], y[1], y[2], y[3]);
                    reinterpret_cast<float4 *>(xnext)[2 * q + 1] = make_float4(y[4], y[5], y[6], y[7]);
                }
            } else if (in_exch_poll) {   // the same sum with every term polled: the residual before this exchange, then the ranks' partials
                unsigned int xc[8];
                if (xparts_prev_poll) { ld_item(xprev, q, xc); while (item_has_sent(xc)) ld_item(xprev, q, xc); }
                else {
                    const uint4 a = __ldcg(reinterpret_cast<const uint4 *>(xprev) + 2 * q), b = __ldcg(reinterpret_cast<const uint4 *>(xprev) + 2 * q + 1);
                    xc[0] = a.x; xc[1] = a.y; xc[2] = a.z; xc[3] = a.w; xc[4] = b.x; xc[5] = b.y; xc[6] = b.z; xc[7] = b.w;
                }
#pragma unroll
                for (int i = 0; i < 8; i++) y[i] = __uint_as_float(xc[i]);
                // the ranks' partials, two at a time in flight, added in rank order: one round trip per pair once they have landed
                // (one rank after the other was tp dependent round trips on the critical path of every exchange)
                for (int rr = 0; rr < tp; rr += 2) {
                    const float *pa = xparts + (size_t)rr * dim, *pb = pa + (rr + 1 < tp ? dim : 0);
                    unsigned int xd[8];
                    ld_item_sys(pa, q, xc);
                    ld_item_sys(pb, q, xd);
                    while (item_has_sent(xc)) { if (poll_ns) __nanosleep(poll_ns); ld_item_sys(pa, q, xc); }
                    while (item_has_sent(xd)) { if (poll_ns) __nanosleep(poll_ns); ld_item_sys(pb, q, xd); }
#pragma unroll
                    for (int i = 0; i < 8; i++) y[i] += __uint_as_float(xc[i]);
                    if (rr + 1 < tp) {
#pragma unroll
                        for (int i = 0; i < 8; i++) y[i] += __uint_as_float(xd[i]);
                    }
                }
                if (xstore) {   // the new residual, polled by the exchange after this one
#pragma unroll
                    for (int i = 0; i < 8; i++) st_poll(xnext, 8 * q + i, y[i]);
                }
                if (r == 0 && ckrow) { ckrow[2] = (unsigned long long)clock64(); ckrow[15] = gtime(); }
            } else if (in_poll) {   // look again until none of my 8 elements reads as the sentinel
                unsigned int xc[8];
                ld_item(px, q, xc);
                while (item_has_sent(xc)) {
                    if (poll_ns) __nanosleep(poll_ns);
                    ld_item(px, q, xc);
                }
#pragma unroll
                for (int i = 0; i < 8; i++) y[i] = __uint_as_float(xc[i]);
                if (r == 0 && ckrow) { ckrow[2] = (unsigned long long)clock64(); ckrow[15] = gtime(); }
            } else {
                const float4 a = __ldcg(reinterpret_cast<const float4 *>(px) + 2 * q), b = __ldcg(reinterpret_cast<const float4 *>(px) + 2 * q + 1);
                y[0] = a.x; y[1] = a.y; y[2] = a.z; y[3] = a.w; y[4] = b.x; y[5] = b.y; y[6] = b.z; y[7] = b.w;
            }
            if (normed) {
                float s4 = 0.f;
                if (r == 0) {
#pragma unroll
                    for (int i = 0; i < 8; i++) { s4 = fmaf(y[i], y[i], s4); y[i] *= wv0[i]; }
                } else {
                    const float4 a = __ldg(reinterpret_cast<const float4 *>(pnw) + 2 * q), b = __ldg(reinterpret_cast<const float4 *>(pnw) + 2 * q + 1);
                    const float w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
                    for (int i = 0; i < 8; i++) { s4 = fmaf(y[i], y[i], s4); y[i] *= w[i]; }
                }
                ss += (double)s4;
            }
        }
        float mx = 0.f;
#pragma unroll
        for (int i = 0; i < 8; i++) mx = fmaxf(mx, fabsf(y[i]));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));   // the two lanes of a 16-element half block sit next to each other
        // Q4_0: max|y| * S in [2^10, 2^11) (the low-nibble operands carry 16 * y * S); Q8_0: [2^14, 2^15)
        int es = (TYPE == NL_Q8_0 ? 268 : 264) - (int)((__float_as_uint(mx) >> 23) & 0xFFu);
        es = es < 27 ? 27 : (es > 227 ? 227 : es);
        const float S = __uint_as_float((uint32_t)es << 23);
        const int b = q >> 2, o = (q & 3) * 8;         // block, offset of my 8 elements inside it
        float bs = 0.f;
        uint8_t *fb = xfrag + (size_t)(b >> 2) * TL_XBG + (size_t)(2 * (b & 3)) * TL_XCOL;
        if constexpr (TYPE == NL_Q8_0) {
            // MMA i takes elements 4i..4i+3: registers 2i = (e0, e2), 2i+1 = (e1, e3); my 8 elements fill registers o/2 .. o/2+3
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; j++) { v[j] = y[j] * S; bs += v[j]; }
            uint32_t h[4], l[4];
#pragma unroll
            for (int k = 0; k < 2; k++) {
                h[2 * k] = pack_h2(v[4 * k], v[4 * k + 2]); h[2 * k + 1] = pack_h2(v[4 * k + 1], v[4 * k + 3]);
                const float2 f0 = __half22float2(*reinterpret_cast<const __half2 *>(&h[2 * k])), f1 = __half22float2(*reinterpret_cast<const __half2 *>(&h[2 * k + 1]));
                l[2 * k] = pack_h2(v[4 * k] - f0.x, v[4 * k + 2] - f0.y); l[2 * k + 1] = pack_h2(v[4 * k + 1] - f1.x, v[4 * k + 3] - f1.y);
            }
            if (store) {
                *reinterpret_cast<uint4 *>(fb + o * 2) = make_uint4(h[0], h[1], h[2], h[3]);          // hi column of this block
                *reinterpret_cast<uint4 *>(fb + TL_XCOL + o * 2) = make_uint4(l[0], l[1], l[2], l[3]);     // lo column
            }
        } else {
            const int pos = o >> 4, ib = ((o & 15) >> 3) * 2;  // low | high nibble half, first of my two word indices
#pragma unroll
            for (int k = 0; k < 2; k++) {
                float v[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    v[j] = y[4 * k + j] * S;
                    bs += v[j];
                    if (pos == 0) v[j] *= 16.f;
                }
                const uint32_t h0 = pack_h2(v[0], v[2]), h1 = pack_h2(v[1], v[3]);
                const float2 f0 = __half22float2(*reinterpret_cast<const __half2 *>(&h0)), f1 = __half22float2(*reinterpret_cast<const __half2 *>(&h1));
                const uint32_t l0 = pack_h2(v[0] - f0.x, v[2] - f0.y), l1 = pack_h2(v[1] - f1.x, v[3] - f1.y);
                const int reg = 4 * (ib + k) + 2 * pos;
                if (store) {
                    *reinterpret_cast<uint2 *>(fb + reg * 4) = make_uint2(h0, h1);         // hi column of this block
                    *reinterpret_cast<uint2 *>(fb + TL_XCOL + reg * 4) = make_uint2(l0, l1);    // lo column
                }
            }
        }
        bs += __shfl_xor_sync(0xffffffffu, bs, 1);
        // per half block: -zero_point * 2^-k * sum(y * S) (Q4_0: 8 * 2^-20, Q8_0: 128 * 2^-24 -- both 2^-17) and 2^k / S, which undoes
        // the operand scaling (k = 20 | 24)
        if (store && (q & 1) == 0) corr[2 * b + (o >> 4)] = make_float2(-7.62939453125e-6f * bs, __uint_as_float((uint32_t)((TYPE == NL_Q8_0 ? 278 : 274) - es) << 23));
    }
    return ss;
}
// Phase input that arrives as a fragment image (producer-side fragments, nl_tile.cuh): copied into shared memory in 16-byte chunks, every
// chunk looked at until none of its words reads as the sentinel (a block group is 32 fragment chunks + 8 half-block records; up to
// four chunks per thread in flight).  Blocks beyond the vector (the last block group's padding) become zero fragments with zero
// scales.  Returns this thread's part of the sum of squares the producers left in the records.
__device__ __forceinline__ bool chunk_has_sent(const uint4 &v) { return max(max(v.x, v.y), max(v.z, v.w)) == TL_SENT; }
__device__ __forceinline__ double input_image_body(const uint8_t *img, int nbg, int nb, uint32_t xfrag_u, uint32_t corr_u, int tid, int poll_ns, unsigned long long *ckrow) {
    double ss = 0.0;
    const int n_chunks = nbg * 40;
    constexpr int BATCH = 4;
    for (int c0 = tid; c0 < n_chunks; c0 += BATCH * TL_CONSUMERS) {
        uint4 v[BATCH];
        const uint8_t *src[BATCH];
        uint32_t dst[BATCH];
#pragma unroll
        for (int j = 0; j < BATCH; j++) {
            const int c = c0 + j * TL_CONSUMERS;
            v[j] = make_uint4(0u, 0u, 0u, 0u); src[j] = nullptr; dst[j] = 0u;
            if (c < n_chunks) {
                const int bg = (int)__umulhi((unsigned)c, 107374183u), k = c - bg * 40;   // c / 40
                const int block = 4 * bg + (k < 32 ? (k >> 3) : ((k - 32) >> 1));
                dst[j] = k < 32 ? xfrag_u + (uint32_t)(bg * TL_XBG + (k >> 2) * TL_XCOL + (k & 3) * 16)
                                : (corr_u + (uint32_t)(bg * 64 + (k - 32) * 8)) | 1u;   // bit 0: a record
                if (block < nb) { src[j] = img + (size_t)bg * TL_IMG_BG + k * 16; v[j] = ld_vol_v4(src[j]); }
            }
        }
#pragma unroll
        for (int j = 0; j < BATCH; j++) {
            if (dst[j] == 0u) continue;
            if (src[j]) {
                while (chunk_has_sent(v[j])) {
                    if (poll_ns) __nanosleep(poll_ns);
                    v[j] = ld_vol_v4(src[j]);
                }
            }
            if (j == 0 && c0 == tid && ckrow) { ckrow[2] = (unsigned long long)clock64(); ckrow[15] = gtime(); }
            if (dst[j] & 1u) {
                asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(dst[j] & ~1u), "r"(v[j].x), "r"(v[j].y) : "memory");
                ss += (double)__uint_as_float(v[j].z);
            } else asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(dst[j]), "r"(v[j].x), "r"(v[j].y), "r"(v[j].z), "r"(v[j].w) : "memory");
        }
    }
    return ss;
}
// The less travelled kinds of phase input: a fragment image (NL_TILE_IMG=1), the tensor-parallel exchanges (polled / behind barriers), a
// plain vector (layer 0's embedding, the one-phase GEMV of nl_matrix).
template <int TYPE, int MODE>
__device__ __noinline__ double other_input(TlShared &sh, const TilePhase &P, uint8_t *xfrag, bool xstore, unsigned long long *ckrow) {
    const int tid = threadIdx.x, nbg = P.nbg, nitem = P.cols >> 3, nitem_pad = nbg * 16;
    float2 *corr = reinterpret_cast<float2 *>(xfrag + TL_XFRAG_BYTES);
    if (P.in_img) return input_image_body(P.in_img, nbg, P.cols >> 5, smem_u32(xfrag), smem_u32(corr), tid, sh.poll_ns, ckrow);
    if (P.in_exch && (MODE == 2 || P.parts)) return input_frags_body<TYPE, 3>(nullptr, P.norm_w, nitem, nitem_pad, xfrag, corr, tid, sh.poll_ns, P.prev, P.next, P.parts, sh.tp, sh.dim, xstore, ckrow, P.prev_poll != 0);
    if (MODE == 0 && P.in_exch) return input_frags_body<TYPE, 2>(nullptr, P.norm_w, nitem, nitem_pad, xfrag, corr, tid, 0, P.prev, P.next, sh.ar_mine + (size_t)P.par * sh.tp * sh.dim, sh.tp, sh.dim, xstore, nullptr);
    if (MODE == 0 && P.in_poll) return input_frags_body<TYPE, 1>(P.x, P.norm_w, nitem, nitem_pad, xfrag, corr, tid, sh.poll_ns, nullptr, nullptr, nullptr, 0, 0, false, ckrow);
    return input_frags_body<TYPE, 0>(P.x, P.norm_w, nitem, nitem_pad, xfrag, corr, tid, 0, nullptr, nullptr, nullptr, 0, 0, false, ckrow);
}
// One GEMV phase on the math warps, after the wait for its input: input vector -> fragments, then the streaming loop.  ONE out-of-line
// function per phase kind (this one and attn_item_tiled) with the conversion and the loop inlined into it: its registers are allocated
// for the phase alone, and the phase loop of the kernel keeps almost nothing alive across the call.  Anything spilled around here goes to
// local memory, i.e. to L2 (12 KB of L1 are left next to the ring): measured 2x on the whole token when the hot loop spilled its B
// fragments.  Everything the phase needs beyond the six scalar arguments is read from shared memory.
template <int TYPE, int MODE>
__device__ __noinline__ int gemv_phase(TlShared &sh, const TilePhase &P, uint8_t *smem, int band, bool first_unit, int it, int p) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t *xfrag = smem + (size_t)TL_SLOTS * TL_SLOT_BYTES;
    float2 *corr = reinterpret_cast<float2 *>(xfrag + TL_XFRAG_BYTES);
    const int nbg = P.nbg;
    const bool normed = P.norm_w != nullptr, in_exch = P.in_exch != 0;
    const int nitem = P.cols >> 3, nitem_pad = nbg * 16;
    const bool xstore = in_exch && first_unit;   // (the CTA that holds the matrix's first unit) stores the new residual
    unsigned long long *ckrow = ((MODE == 0 || NL_TL_FINE_TRACE) && sh.trace2 && tid == 0) ? sh.trace2 + ((size_t)blockIdx.x * sh.n_phases + p) * 16 : nullptr;   // (slot 15: coarse trace too)
    unsigned long long *trrow = ((MODE == 0 || NL_TL_FINE_TRACE) && sh.trace && tid == 0) ? sh.trace + ((size_t)blockIdx.x * sh.n_phases + p) * 8 : nullptr;
#if NL_TL_FINE_TRACE
    if (ckrow) ckrow[1] = (unsigned long long)clock64();
#endif
    // the single-GPU hot path (a polled fp32 vector) is inlined; every other kind of input goes through one out-of-line function, so that
    // its code does not weigh on this function's register allocation (the same source has allocated differently across unrelated edits)
    double ss;
    if (MODE == 1 || (MODE == 0 && P.in_poll && !P.in_img && !in_exch)) ss = input_frags_body<TYPE, 1>(P.x, P.norm_w, nitem, nitem_pad, xfrag, corr, tid, sh.poll_ns, nullptr, nullptr, nullptr, 0, 0, false, ckrow);
    else ss = other_input<TYPE, MODE>(sh, P, xfrag, xstore, ckrow);
    if (trrow && ckrow && (P.in_poll || P.in_img || (in_exch && P.parts))) trrow[1] = ckrow[15];   // "first item valid" (globaltimer)
#if NL_TL_FINE_TRACE
    if (ckrow) ckrow[3] = (unsigned long long)clock64();
    if (tid == TL_CONSUMERS - 32 && sh.trace2) sh.trace2[((size_t)blockIdx.x * sh.n_phases + p) * 16 + 6] = (unsigned long long)clock64();
#endif
    if (normed) {
        // every thread's part is an fp32 sum of 8 squares (x3 at most); the warp's 32 parts are folded in fp32 too (five double-precision
        // shuffle rounds were ~300 cycles on the critical path of every normed phase), the 16 warp sums in float64 by the finishing warp
        const float sw = warp_sum((float)ss);
        if (lane == 0) sh.ss_red[warp] = (double)sw;
    }
    tl_bar<TL_CONSUMERS>();   // fragments complete; the finishing warp turns ss_red into the RMSNorm scale once the first slot is consumed
    if (trrow) trrow[2] = gtime();
#if NL_TL_FINE_TRACE
    if (ckrow) ckrow[4] = (unsigned long long)clock64();
#endif
    if (xstore && !P.parts) __threadfence();
    if (TileCfg<TYPE>::TPW == 2 && (nbg & 1) == 0) it = stream_band<TYPE, true>(band, nbg, P.nbg_magic, it, smem_u32(smem), smem_u32(&sh));
    else it = stream_band<TYPE, false>(band, nbg, P.nbg_magic, it, smem_u32(smem), smem_u32(&sh));
    if (trrow) trrow[3] = gtime();
#if NL_TL_FINE_TRACE
    if (ckrow) ckrow[5] = (unsigned long long)clock64();
    if (tid == TL_CONSUMERS - 32 && sh.trace2) sh.trace2[((size_t)blockIdx.x * sh.n_phases + p) * 16 + 7] = (unsigned long long)clock64();
#endif
    return it;
}

template <int TYPE, int MODE>
// One CTA of 576 threads = 18 warps per SM.  Registers are handed out per warp, to warps in multiples of four: 20 warps' worth has to fit
// the 64 K registers, i.e. 96 per thread (that is where __launch_bounds__(576, 1) lands; a cap of 112 compiles without the spills of
// the phase loop but cannot be launched: "too many blocks in cooperative launch").
__global__ void __launch_bounds__(TL_THREADS, 1) decode_tiled_kernel(const TileArgs A) {
    constexpr int TILE = TileCfg<TYPE>::TILE, TPW = TileCfg<TYPE>::TPW, TS = TPW * TL_CW;
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ TlShared sh;
    uint8_t *ring = smem;
    uint8_t *xfrag = smem + (size_t)TL_SLOTS * TL_SLOT_BYTES;
    // (behind the fragments: per block {zero-point correction, 2^20 / S_b}, see gemv_phase)
    AttnT &att = *reinterpret_cast<AttnT *>(xfrag);   // the attention phase has no GEMV input: same bytes

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int G = gridDim.x;
    const unsigned int epoch = A.epoch ? __ldg(A.epoch) : 0u;
    const unsigned int flag_base = epoch * (unsigned)(A.n_phases + 1);
    // MODE 1: the instantiation for the default single-GPU run (polled fp32 vectors, no fragment images, no tracing, no forensics): every
    // tensor-parallel, barrier, image and trace branch below folds away.  MODE 2: the same for the polled tensor-parallel run (exchanges
    // and the two local fragment images stay, barriers / tracing / forensics go).  MODE 0: everything.  Measured on big (10 layers):
    // 446 -> 401 us per token from the lean instantiation alone -- the phase boundaries walk through a lot of code.
    const bool poll = MODE != 0 || A.poll != 0;
    const bool tpar = MODE == 2 || (MODE == 0 && A.tp > 1);
    const int dbg = MODE != 0 ? 0 : A.dbg;
    if (warp == 1) {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(&A.at);
        for (int i = lane; i < (int)(sizeof(MegaAttn) / 4); i += 32) reinterpret_cast<uint32_t *>(&sh.at)[i] = src[i];
    }
    if (tid == 0) {
        sh.trace = (MODE != 0 && !NL_TL_FINE_TRACE) ? nullptr : A.trace; sh.trace2 = (MODE != 0 && !NL_TL_FINE_TRACE) ? nullptr : A.trace2; sh.n_phases = A.n_phases; sh.poll_ns = A.poll_ns; sh.tp = A.tp; sh.dim = A.dim;
        sh.ar_mine = tpar ? reinterpret_cast<const float *>(A.peers.win[A.rank] + A.ar_off) : nullptr;
        for (int s = 0; s < TL_SLOTS; s++) { mbar_init(&sh.full_bar[s], 1); mbar_init(&sh.empty_bar[s], TL_CW); mbar_init(&sh.free_bar[s], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == TL_CW) {
        // ===================== copy warp: streams this CTA's band of every GEMV phase, in phase order =====================
        if (lane != 0) return;
        uint64_t policy;   // weights are read once per token: keep them from evicting activations / KV / norm weights out of L2
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
        // Cursor over this CTA's slots in stream order: (phase, first tile of the slot inside the band).  The copies use one; with
        // A.l2pf > 0 a second one runs TL_SLOTS + l2pf slots ahead and asks the L2 for those bytes (cp.async.bulk.prefetch.L2): while the
        // math warps sit at a phase boundary the ring is full and the copies stop, but HBM keeps streaming into the 126 MB L2, and the
        // slots after the boundary are refilled from there.
        struct Cur { int p, c0, band; const uint8_t *src; };
        auto seek = [&](Cur &c) {   // first GEMV phase at or after c.p in which this CTA has tiles (c.p = n_phases: exhausted)
            for (; c.p < A.n_phases; c.p++) {
                const TilePhase *P = A.phases + c.p;
                if (__ldg(&P->kind) != PH_GEMV) continue;
                const int nbg = __ldg(&P->nbg), urg = __ldg(&P->unit_rg);
                int u0, u1;
                band_of(__ldg(&P->units), blockIdx.x, G, A.g_magic, u0, u1);
                c.band = (u1 - u0) * urg * nbg;
                if (c.band == 0) continue;
                c.src = ldg_ptr(&P->tiles) + (size_t)u0 * urg * nbg * TILE;
                c.c0 = 0;
                return;
            }
        };
        auto advance = [&](Cur &c) { c.c0 += TS; if (c.c0 >= c.band) { c.p++; seek(c); } };
        Cur cc{0, 0, 0, nullptr}, pc{0, 0, 0, nullptr};
        seek(cc);
        int pf_it = 0;
        if (A.l2pf > 0) { seek(pc); for (; pf_it < TL_SLOTS && pc.p < A.n_phases; pf_it++) advance(pc); }   // the ring itself needs no prefetch
        for (int it = 0; cc.p < A.n_phases; it++, advance(cc)) {
            const int slot = it % TL_SLOTS;
            for (; A.l2pf > 0 && pf_it < it + TL_SLOTS + A.l2pf && pc.p < A.n_phases; pf_it++, advance(pc)) {
                const uint32_t pbytes = (uint32_t)min(TS, pc.band - pc.c0) * TILE;
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(pc.src + (size_t)pc.c0 * TILE), "r"(pbytes) : "memory");
            }
            if (it >= TL_SLOTS) mbar_wait_parked(&sh.free_bar[slot], ((it / TL_SLOTS) - 1) & 1);
            // at most `inflight` copies on the wire: bytes requested but not yet landed are queue in front of every other
            // request of this SM (barrier polls, the phase input, KV rows), and ~2 slots already cover latency x bandwidth
            if (it >= A.inflight) mbar_wait_parked(&sh.full_bar[(it - A.inflight) % TL_SLOTS], ((it - A.inflight) / TL_SLOTS) & 1);
            const uint32_t bytes = (uint32_t)min(TS, cc.band - cc.c0) * TILE;
            if (dbg == 1 || dbg >= 3) { mbar_arrive(&sh.full_bar[slot]); continue; }   // (forensics: no copy)
            mbar_expect_tx(&sh.full_bar[slot], bytes);
            bulk_g2s_hint(ring + (size_t)slot * TL_SLOT_BYTES, cc.src + (size_t)(dbg == 2 ? cc.c0 % (2 * TS) : cc.c0) * TILE, bytes, &sh.full_bar[slot], policy);
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
            const float *resid_src = ldg_ptr(&P->resid);
            uint8_t *out_img = MODE == 1 ? nullptr : ldg_ptr(&P->out_img);          // producer-side fragments of the next GEMV's input (nl_tile.cuh)
            const float *out_nw = MODE == 1 ? nullptr : ldg_ptr(&P->out_nw);
            const int out_poll = __ldg(&P->out_poll), resid_poll = __ldg(&P->resid_poll);   // polled vectors (see st_poll)
            const int exch_out = MODE == 1 ? 0 : __ldg(&P->exch_out), par = MODE != 0 ? 0 : __ldg(&P->par), cross = MODE != 0 ? 0 : __ldg(&P->cross);
            const unsigned long long exch_off = MODE == 1 ? 0ull : __ldg(&P->exch_off);
            const bool want_logits = MODE == 1 || !A.lg_want || __ldg(A.lg_want) != 0;
            const bool to_peers_logits = tpar && p == A.n_phases - 1;
            const bool normed = ldg_ptr(&P->norm_w) != nullptr;
            const int cols_p = __ldg(&P->cols);
            int u0, u1;
            band_of(__ldg(&P->units), blockIdx.x, G, A.g_magic, u0, u1);
            const int band = (u1 - u0) * urg * nbg;
            const int rg0 = u0 * urg;
            float racc = 0.f, gate = 0.f, post = 1.f;
            if (lane == 0) TL_CK(p, 8);
            float best = -INFINITY;
            int best_i = 0x7fffffff;
            // ---- the lean loop: local outputs only (every phase of a single GPU run, the column-split phases of a tensor-parallel one).
            // One warp walks every slot of the CTA, so each instruction here is latency on the slot recycling path and, for the last slot of a
            // phase, on the token's critical path: row groups are tracked incrementally (no divides, no searches), whatever the completing
            // row group needs from memory (bias, residual) is requested before the wait, and the common slot -- 32 tiles of one row group --
            // is eight fixed loads, one tree, one shuffle.
            if (!exch_out && !to_peers_logits && !out_img && dbg != 3) {
                int rem = nbg;          // tiles of the current row group still to come
                int rg = rg0;           // the current row group (of the whole matrix)
                const float *rb0 = &sh.red[0][0][0][row];
                for (int c0 = 0; c0 < band; c0 += TS, it++) {
                    const int slot = it % TL_SLOTS;
                    const int n = min(TS, band - c0);
                    const bool done0 = rem <= n;                        // the current row group ends inside this slot
                    const int r0 = (epi == TEPI_SWIGLU ? (rg >> 1) : rg) * 16 + row;
                    float bv = 0.f, resid = 0.f;
                    bool have_resid = false;
                    if (done0 && r0 < rows && epi != TEPI_SWIGLU) {
                        if (bias) bv = __ldg(bias + r0);
                        if (epi == TEPI_RESID && (resid_poll || c0 > 0)) { resid = resid_poll ? ld_poll(resid_src, r0) : __ldcg(resid_src + r0); have_resid = true; }
                    }
#if NL_TL_FIN_SPIN
                    mbar_wait_u(smem_u32(&sh.empty_bar[slot]), (uint32_t)(it / TL_SLOTS) & 1u);
#else
                    mbar_wait_parked(&sh.empty_bar[slot], (it / TL_SLOTS) & 1);
#endif
                    if (lane == 0 && c0 + n == band) { TL_TRACE(p, 5); TL_CK(p, 9); }   // last slot consumed by every math warp
                    if (c0 == 0 && normed) {   // RMSNormInto, go/quant.go:597-607 (see the general loop below)
                        double s2 = sh.ss_red[lane & (TL_CW - 1)];   // 16 warp sums, folded in float64 in a fixed butterfly order (both half-warps alike)
#pragma unroll
                        for (int o = 1; o < TL_CW; o <<= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
                        post = rsqrtf((float)s2 * (1.0f / (float)cols_p) + A.eps);
                    }
                    const float *rb = rb0 + slot * (TL_CW * 2 * 16);
                    int a = 0;                                           // first tile of the slot that belongs to the current row group
                    bool first = true;
                    do {
                        const int take = min(rem, n - a);               // tiles [a, a + take) of this slot belong to row group rg
                        float pv[TL_CW / 2];
                        if (take == TS) {                                // the whole slot is one row group: every warp's entry 0
#pragma unroll
                            for (int i = 0; i < TL_CW / 2; i++) pv[i] = rb[(half + 2 * i) * 32];
                        } else {
                            const int wa = a / TPW, wb = (a + take - 1) / TPW;
#pragma unroll
                            for (int i = 0; i < TL_CW / 2; i++) {
                                const int w = wa + half + 2 * i;         // entry 1 when the warp's first tile still belonged to the group before
                                pv[i] = (w <= wb) ? rb[w * 32 + ((w == wa && TPW * w < a) ? 16 : 0)] : 0.f;
                            }
                        }
                        float sum = ((pv[0] + pv[1]) + (pv[2] + pv[3])) + ((pv[4] + pv[5]) + (pv[6] + pv[7]));
                        sum += __shfl_xor_sync(0xffffffffu, sum, 16);
                        racc += sum;
                        rem -= take; a += take;
                        if (rem == 0) {                                  // the row group is complete: publish its 16 rows
                            float v = racc * post;
                            racc = 0.f; rem = nbg;
                            if (epi == TEPI_SWIGLU) {
                                if ((rg & 1) == 0) gate = v;
                                else {
                                    const int r = (rg >> 1) * 16 + row;
                                    if (half == 0 && r < rows) { const float hv = silu_f(gate) * v; if (out_poll) st_poll(out, r, hv); else out[r] = hv; }   // SiLU(gate)*up, go/model.go:604-606
                                }
                            } else {
                                const int r = rg * 16 + row;
                                if (half == 0 && r < rows) {
                                    if (first) v += bv; else if (bias) v += __ldg(bias + r);
                                    if (epi == TEPI_RESID) v += (first && have_resid) ? resid : (resid_poll ? ld_poll(resid_src, r) : __ldcg(resid_src + r));   // X += W.x, go/model.go:592-594, :610-612
                                    if (out_poll) st_poll(out, r, v); else out[r] = v;
                                    if (v > best || (v == best && r < best_i)) { best = v; best_i = r; }   // first maximum, go/main.go:400-408 (LM head)
                                }
                            }
                            rg++;
                        }
                        first = false;
                    } while (a < n);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&sh.free_bar[slot]);
                }
            } else
            for (int c0 = 0; c0 < band; c0 += TS, it++) {
                const int slot = it % TL_SLOTS;
                const int c1 = min(c0 + TS, band);
                const int q_first = rg_of(c0, nbg, magic), q_last = rg_of(c1 - 1, nbg, magic);
                // the residual of a row group that completes in this slot is fetched before we block on the math warps (not for the first
                // slot of a phase: only a consumed slot proves that this CTA is past the grid barrier that orders the residual's writers)
                float resid = 0.f, nwv = 1.f;
                int q_done = -1;
                if ((epi == TEPI_RESID || out_nw) && !exch_out) {
                    for (int q = q_first; q <= q_last; q++) if ((q + 1) * nbg <= c1) { q_done = q; break; }
                    if (q_done >= 0) {
                        const int r = (rg0 + q_done) * 16 + row;
                        if (out_nw && r < rows) nwv = __ldg(out_nw + r);   // static data
                        if (epi == TEPI_RESID && c0 > 0 && r < rows) resid = resid_poll ? ld_poll(resid_src, r) : __ldcg(resid_src + r);
                        else if (epi == TEPI_RESID) q_done = -2 - q_done;   // first slot: the residual is fetched after the wait (-2 - q: nwv is valid)
                    }
                }
#if NL_TL_FIN_SPIN   // (variant: the finishing warp re-issues try_wait instead of parking -- wakes sooner, costs issue slots)
                mbar_wait_u(smem_u32(&sh.empty_bar[slot]), (uint32_t)(it / TL_SLOTS) & 1u);
#else
                mbar_wait_parked(&sh.empty_bar[slot], (it / TL_SLOTS) & 1);
#endif
                if (lane == 0 && c1 == band) { TL_TRACE(p, 5); TL_CK(p, 9); }   // last slot consumed by every math warp
                if (c0 == 0 && normed) {   // RMSNormInto, go/quant.go:597-607: inv = 1 / sqrt(ss / n + eps) from the float64 sum of squares
                    double s2 = warp_sum_d(lane < TL_CW ? sh.ss_red[lane] : 0.0);   // fixed butterfly order: deterministic
                    // float64 sum like the reference; the final 1/sqrt in fp32 (within 1 ulp of the reference's float32(1/sqrt(float64)))
                    post = rsqrtf((float)s2 * (1.0f / (float)cols_p) + A.eps);
                }
                for (int q = q_first; q <= q_last; q++) {
                    if (dbg == 3 && (q + 1) * nbg > c1) continue;   // (forensics)
                    const int a = max(q * nbg, c0) - c0, b = min((q + 1) * nbg, c1) - c0;   // tiles [a, b) of this slot belong to row group q
                    const int wa = a / TPW, wb = (b - 1) / TPW;   // math warp w owns slot tiles [TPW * w, TPW * w + TPW)
                    // warp w dropped this row group's sums into entry 0, unless its first tile still belonged to the previous group
                    // (then: entry 1); independent loads first, one fixed summation tree after (deterministic)
                    float pv[TL_CW / 2];
                    const float *rb = &sh.red[slot][0][0][row];   // + 32 floats per warp, + 16 for entry 1
                    if (a == 0 && b == TS) {   // the common case: the whole slot is one row group, every warp's entry 0
#pragma unroll
                        for (int i = 0; i < TL_CW / 2; i++) pv[i] = rb[(half + 2 * i) * 32];
                    } else {
#pragma unroll
                        for (int i = 0; i < TL_CW / 2; i++) {
                            const int w = wa + half + 2 * i;
                            pv[i] = (w <= wb) ? rb[w * 32 + ((w == wa && TPW * w < a) ? 16 : 0)] : 0.f;
                        }
                    }
                    float s = ((pv[0] + pv[1]) + (pv[2] + pv[3])) + ((pv[4] + pv[5]) + (pv[6] + pv[7]));
                    s += __shfl_xor_sync(0xffffffffu, s, 16);
                    racc += s;
                    if ((q + 1) * nbg <= c1) {   // last tile of the row group is in this slot: publish its 16 rows (both half-warps hold them)
                        const int rg = rg0 + q;
                        float v = racc * post;
                        racc = 0.f;
                        if (epi == TEPI_SWIGLU) {
                            if ((rg & 1) == 0) gate = v;
                            else {
                                const int r = (rg >> 1) * 16 + row;
                                const float hv = r < rows ? silu_f(gate) * v : 0.f;                  // SiLU(gate)*up, go/model.go:604-606
                                if (out && half == 0 && r < rows) { if (out_poll) st_poll(out, r, hv); else out[r] = hv; }
                                if (out_img) publish_half_block<TYPE>(out_img, (rg >> 1) * 16, row, half, hv, 0.f);
                            }
                        } else {
                            const int r = rg * 16 + row;
                            const bool ok = r < rows;
                            if (bias && ok) v += __ldg(bias + r);
                            if (exch_out) {          // this rank's partial of a row-split product -> slot `rank` on every rank (NVLink stores)
                                if (half == 0 && ok) {
                                    if (poll) {      // polled slots of this token parity's arena: the consumers look at the data itself
                                        for (int rr = 0; rr < A.tp; rr++)
                                            st_poll_sys(reinterpret_cast<float *>(A.peers.win[rr] + exch_off) + (size_t)A.rank * A.dim + r, v);
                                    } else {
                                        for (int rr = 0; rr < A.tp; rr++)
                                            reinterpret_cast<float *>(A.peers.win[rr] + A.ar_off)[((size_t)par * A.tp + A.rank) * A.dim + r] = v;
                                    }
                                }
                            } else if (to_peers_logits) {   // vocab-split LM head: my rows of the full logits vector on every rank (when somebody reads them)
                                if (half == 0 && ok && want_logits)
                                    for (int rr = 0; rr < A.tp; rr++) reinterpret_cast<float *>(A.peers.win[rr] + A.lg_off)[(size_t)A.rank * A.lvocab + r] = v;
                            } else {
                                if (epi == TEPI_RESID && ok) v += (q == q_done) ? resid : (resid_poll ? ld_poll(resid_src, r) : __ldcg(resid_src + r));   // X += W.x, go/model.go:592-594, :610-612
                                if (out && half == 0 && ok) { if (out_poll) st_poll(out, r, v); else out[r] = v; }
                                if (out_img) {       // the next GEMV reads RMSNorm(out; out_nw): y = v o w here, 1/rms on its finished sums
                                    const float w = !out_nw ? 1.f : ((q == q_done || q == -2 - q_done) ? nwv : (ok ? __ldg(out_nw + r) : 1.f));
                                    publish_half_block<TYPE>(out_img, rg * 16, row, half, ok ? v * w : 0.f, ok ? v * v : 0.f);
                                }
                            }
                            if (half == 0 && ok) {
                                const int gr = tpar ? A.rank * A.lvocab + r : r;   // (only meaningful in the LM-head phase)
                                if (v > best || (v == best && gr < best_i)) { best = v; best_i = gr; }   // first maximum, go/main.go:400-408
                            }
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&sh.free_bar[slot]);
            }
            if (A.amax && p == A.n_phases - 1) {   // LM head: this CTA's (maximum, first index) for the greedy step that follows
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
                    if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
                }
                if (lane == 0) {
                    if (tpar && poll) {   // polled pairs: the last thing a rank publishes; the logits it stored before them are ordered by the fence
                        if (want_logits) __threadfence_system();
                        for (int rr = 0; rr < A.tp; rr++)
                            asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1, %2};" ::"l"(reinterpret_cast<float2 *>(A.peers.win[rr] + A.amax_off) + A.rank * gridDim.x + blockIdx.x),
                                         "r"(__float_as_uint(best) == TL_SENT ? 0x7FFFFFFFu : __float_as_uint(best)), "r"((unsigned)best_i) : "memory");
                    } else if (tpar) {
                        for (int rr = 0; rr < A.tp; rr++)
                            reinterpret_cast<float2 *>(A.peers.win[rr] + A.amax_off)[A.rank * gridDim.x + blockIdx.x] = make_float2(best, __int_as_float(best_i));
                    } else A.amax[blockIdx.x] = make_float2(best, __int_as_float(best_i));
                }
            }
            __syncwarp();
            if (lane == 0) { TL_CK(p, 10); TL_TRACE(p, 4); if (!poll) tl_arrive(A, p, cross != 0); TL_TRACE(p, 6); }   // polled outputs need no arrival
            __syncwarp();
        }
        return;
    }

    // ===================== math warps =====================
    int it = 0;
    if (tid == 0) {   // the attention phases' shape (the same in every layer)
        const int pos = A.at.pos ? *A.at.pos : 0, n = pos + 1;   // (a one-phase GEMV launch has no attention state)
        int nse = (n + max(A.att_chunk, 1) - 1) / max(A.att_chunk, 1);   // one prefetched pass (att_chunk <= 96 positions) per split while the splits last
        nse = nse < 1 ? 1 : (nse > A.at.nsplit ? A.at.nsplit : nse);
        const int hpi = A.at.n_kv_heads > 0 ? attn_hpi(A.at, nse, G, A.att_hpi) : 1;
        sh.a_pos = pos; sh.a_nse = nse; sh.a_hpi = hpi; sh.a_items = A.at.n_kv_heads > 0 ? (A.at.n_heads / hpi) * nse : 0;
        if (A.at.n_kv_heads > 0) sh.a_item0 = attn_locate(A.at, blockIdx.x, n, nse, hpi);
    }
    if (warp == 1) {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(&A.phases[0]);
        uint32_t *dst = reinterpret_cast<uint32_t *>(&sh.ph[0]);
        for (int i = lane; i < (int)(sizeof(TilePhase) / 4); i += 32) dst[i] = __ldg(src + i);
    }
    for (int p = 0; p < A.n_phases; p++) {
        if (warp == 1 && p + 1 < A.n_phases) {   // next descriptor while this phase runs: no L2 round trip after the barrier
            const uint32_t *src = reinterpret_cast<const uint32_t *>(&A.phases[p + 1]);
            uint32_t *dst = reinterpret_cast<uint32_t *>(&sh.ph[(p + 1) % 3]);
            for (int i = lane; i < (int)(sizeof(TilePhase) / 4); i += 32) dst[i] = __ldg(src + i);
        }
        if (p == 0) tl_bar<TL_CONSUMERS>();
        // Three descriptor slots: warp 1 overwrites slot (p + 1) % 3, last read in phase p - 2, and every warp has left that phase (each
        // iteration has a block barrier in front of its streaming loop).  So the fields are read from shared memory where they are used
        // instead of being carried in registers across the phase (they were spilled to local memory: an L2 trip each, see input_frags).
        const TilePhase &P = sh.ph[p % 3];
        const int kind = P.kind, nbg = P.nbg;
        int u0, u1;
        band_of(P.units, blockIdx.x, G, A.g_magic, u0, u1);
        const int band = (kind == PH_GEMV) ? (u1 - u0) * P.unit_rg * nbg : 0;
        if (kind == PH_ATTN) {
            tl_bar<TL_CONSUMERS>();   // every math warp is done with the previous phase's fragments: the buffer becomes attention scratch
            const int pos = sh.a_pos, nse = sh.a_nse, hpi = sh.a_hpi, n_items = sh.a_items;
            const int kvd = A.at.n_kv_heads * 64;
            bool pre = false;
            if ((int)blockIdx.x < n_items) {   // cached K/V rows of my first item while q / k / v are still being produced
                const int t_begin = sh.a_item0.t_begin;
                attn_fetch(A.at.kcache + (size_t)P.layer * A.at.seq_len * kvd, A.at.vcache + (size_t)P.layer * A.at.seq_len * kvd, kvd, sh.a_item0.kvh, t_begin,
                           min(TA_CH, sh.a_item0.t_end - t_begin), pos, att, tid);
                pre = true;
            }
            if (tid == 0) { TL_CK(p, 0); TL_TRACE(p, 0); }
            if (!poll) {   // (barrier mode: the wait for q / k / v, handed to the block)
                if (tid == 0) { tl_wait(A, p - 1, P.wait_cross != 0, (unsigned)G, epoch); TL_TRACE(p, 1); }
                tl_bar<TL_CONSUMERS>();
            }
            if (tid == 0) TL_CK(p, 1);
            for (int item = blockIdx.x; item < n_items; item += G) {
                attn_item_tiled<TYPE, MODE>(sh, att, item, nse, hpi, pos, p, pre, poll, flag_base + (unsigned)p + 1u);
                pre = false;
            }
            if (poll) { if (tid == 0) { TL_TRACE(p, 3); TL_TRACE(p, 4); } continue; }   // (attn_item_tiled ends on a block barrier; the KV rows are for later tokens)
            __threadfence();
            tl_bar<TL_CONSUMERS>();   // every thread has issued its output stores
            if (tid == 0) { TL_TRACE(p, 3); tl_arrive(A, p, false); TL_TRACE(p, 4); }
            continue;
        }

        // ---- prologue: phase input -> fp16 hi/lo B fragments in shared memory ----
        if (tid == 0) TL_CK(p, 0);
        if (P.in_poll && tid == 0) TL_TRACE(p, 0);
        if (p > 0) {
            if (!poll && tid == 0) { TL_TRACE(p, 0); tl_wait(A, p - 1, P.wait_cross != 0, (unsigned)G, epoch); TL_TRACE(p, 1); }
            // also: every math warp is done with the previous phase's fragments (a polled element can be complete while a slower warp of
            // this CTA still streams the phase before), and warp 1's descriptor prefetch is ordered against its readers
            tl_bar<TL_CONSUMERS>();
        }
        if (band == 0) continue;   // nothing of this matrix lands here (the finishing warp has arrived for us)
        it = gemv_phase<TYPE, MODE>(sh, P, smem, band, u0 == 0, it, p);
    }
    // tensor parallel: the kernel may only complete when every rank's logits shard and argmax pairs have landed in this window
    if (tpar && poll) {   // (polled) every CTA of every rank has published its argmax pair, after its logits rows
        const uint2 *pairs = reinterpret_cast<const uint2 *>(A.peers.win[A.rank] + A.amax_off);
        for (int i = tid; i < A.tp * G; i += TL_CONSUMERS) {
            uint2 v;
            do { asm volatile("ld.relaxed.sys.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(pairs + i) : "memory"); } while (v.x == TL_SENT || v.y == TL_SENT);
        }
    } else if (tpar && tid == 0) tl_wait(A, A.n_phases - 1, true, (unsigned)G, epoch);
}

template <int TYPE, int MODE>
static int launch_tiled_t(const TileArgs &a_in, int grid, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(decode_tiled_kernel<TYPE, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TL_DYN_SMEM) != cudaSuccess) return -2;
        configured = true;
    }
    TileArgs a = a_in;
    a.g_magic = grid <= 1 ? 0u : (unsigned int)(((1ull << 32) + (unsigned)grid - 1) / (unsigned)grid);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(TL_THREADS); cfg.dynamicSmemBytes = TL_DYN_SMEM; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative;  // all CTAs must be co-resident: they spin on each other
    at[0].val.cooperative = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, decode_tiled_kernel<TYPE, MODE>, a) == cudaSuccess ? 0 : -2;
}
// the lean instantiations when the phase list allows them (TileArgs::slim: 1 = single GPU, every GEMV input a polled fp32 vector;
// 2 = polled tensor parallel) and nothing asks for the branches they fold away
int launch_tiled(int type, const TileArgs &a, int grid, cudaStream_t st) {
    const bool lean = a.slim && a.poll && (NL_TL_FINE_TRACE || (!a.trace && !a.trace2)) && !a.dbg && !getenv("NL_TILE_NO_SLIM");
    const int mode = !lean ? 0 : (a.slim == 1 && a.tp <= 1) ? 1 : (a.slim == 2 && a.tp > 1) ? 2 : 0;
    if (type == NL_Q8_0) return mode == 1 ? launch_tiled_t<NL_Q8_0, 1>(a, grid, st) : mode == 2 ? launch_tiled_t<NL_Q8_0, 2>(a, grid, st) : launch_tiled_t<NL_Q8_0, 0>(a, grid, st);
    return mode == 1 ? launch_tiled_t<NL_Q4_0, 1>(a, grid, st) : mode == 2 ? launch_tiled_t<NL_Q4_0, 2>(a, grid, st) : launch_tiled_t<NL_Q4_0, 0>(a, grid, st);
}

int launch_tile_repack(int type, const uint8_t *qs, const __half *d, int rows, int nb, uint8_t *tiles, int nbg, int rg_off, int rg_stride, cudaStream_t st) {
    const int n_rg_src = (rows + 15) / 16;
    const long long threads = (long long)n_rg_src * nbg * 32;
    const unsigned grid = (unsigned)((threads + 255) / 256);
    if (type == NL_Q8_0) tile_q8_0_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const uint4 *>(qs), d, rows, nb, tiles, nbg, rg_off, rg_stride, n_rg_src);
    else tile_q4_0_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const uint4 *>(qs), d, rows, nb, tiles, nbg, rg_off, rg_stride, n_rg_src);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

}  // namespace nl
