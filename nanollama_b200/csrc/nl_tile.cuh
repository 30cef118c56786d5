// nl_tile.cuh — batch-1 decode on tensor cores: persistent per-token kernel over MMA-fragment-tiled Q4_0 weights (sm_100a).
//
// Replaces (*LlamaModel).Forward, go/model.go:490-620, with MatMulQ4_0 (go/quant.go:45-94) as the hot loop.
//
// Why tensor cores for a GEMV: at HBM speed each SM sub-partition must retire ~12 Q4_0 weights per cycle; turning
// every nibble into an fp32 operand on the ALU pipe (one PRMT each) caps a CUDA-core kernel near half the HBM roofline
// (profiles/r01_gemv_stream_full.md).  Here a nibble becomes an MMA operand with 0.625 ALU ops: masked in place it IS an
// fp16 subnormal (n * 2^-24, or 16n * 2^-24 one nibble higher), so `mma.sync.m16n8k16` (A = 16 weight rows x 16 nibbles,
// straight from the packed words; B = the activation vector) does the multiply-adds.  Products are exact, accumulation is
// fp32.  The activation x (fp32) enters as an exact-to-2^-22 pair of fp16 terms (hi + lo) in two B columns; four quant
// blocks share one accumulator fragment by giving each block its own column pair, so the per-block scale d (applied AFTER
// the block dot like the reference, quant.go:88) costs one FMA per (row, block) on a fully used warp.
//
// Weight layout ("tiles"): 16 rows x 4 blocks (128 columns) = 1152 contiguous bytes, already in fragment order:
//     [   0, 512)  lane (g,t) -> 16 B: nibbles of block (row g,   blk t)      g = lane >> 2, t = lane & 3
//     [ 512,1024)  lane (g,t) -> 16 B: nibbles of block (row g+8, blk t)
//     [1024,1152)  lane (g,t) ->  4 B: fp16 d(row g, blk t), fp16 d(row g+8, blk t)
// Same bytes as the GGUF tensor (18 B per block), tiles of a 16-row group contiguous, row groups consecutive: a CTA's share
// of a matrix is ONE byte range, moved through a shared-memory ring with cp.async.bulk; every LDS is a conflict-free 512-B
// warp access.
//
// Kernel structure (one CTA per SM, cooperative): 16 math warps + 1 copy warp + 1 finishing warp walk the phase list
//     per layer: QKV (RMSNorm fused) | attention | O (+residual) | gate/up (RMSNorm fused, SiLU*up) | down (+residual);  LM head
// Phase boundaries: on one GPU every activation vector of a token is written once into a sentinel-filled arena and the consumers
// poll the data itself (no grid barrier, no fence; "polled activations" in nl_tile.cu); under tensor parallelism (and with
// NL_TILE_POLL=0) release / acquire grid barriers, one counter per phase.  The copy warp never waits for a phase boundary
// (weights do not depend on activations), so the ring keeps HBM busy across them.  Per ring slot (32 tiles) each math warp owns
// two tiles and drops per-row partial sums into shared memory; the finishing warp adds them in a fixed order (deterministic),
// keeps the running sum of a row group across slots, applies bias / residual / SiLU*up and publishes the outputs.
// Shared memory is 217 KB of the SM's 228, so ~12 KB of L1 remain: whatever ptxas spills goes to L2.  Hence one out-of-line
// function per phase kind, one set of B fragments in the streaming loop, descriptors read from shared memory where they are used.
#pragma once
#include <stdlib.h>

#include "nl_common.cuh"
#include "nl_stream.cuh"  // PTX wrappers
#include "nl_tp.cuh"      // TpPeers: peer windows of a tensor-parallel group

namespace nl {

constexpr int MG_MAX_GROUP = 8;                    // q heads per kv head handled by one attention item
constexpr int MG_MAX_SPLIT = 16;                   // splits of the context per kv head (flash-decoding partials)
enum { PH_GEMV = 0, PH_ATTN = 1 };

// what the attention phase needs besides its descriptor (go/model.go:530-587)
struct MegaAttn {
    const float *q, *k, *v;      // q is the base of this layer's q | k | v vector; k / v only carry element offsets from it
    float *kcache, *vcache;      // [L][S][kvd]
    const float *cos_t, *sin_t;  // [S][hd/2]
    const int32_t *pos;          // device scalar
    float *part_acc;             // [H][nsplit][hd]   un-normalised partial outputs (flagged pairs)
    float *part_ml;              // [H][nsplit][2]    (running max, sum of exp)
    float *out;
    int n_heads, n_kv_heads, seq_len, qk_norm, conj, nsplit;
    float eps, scale;
};

__device__ __forceinline__ unsigned int ld_acquire(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Grid barrier (barrier mode only): every CTA adds 1 to the phase's counter after its last output of that phase (release); one thread
// per CTA polls the counter (acquire).  Counters are zeroed by a memset node in front of the kernel.
__device__ __forceinline__ void phase_arrive(unsigned int *bar, int p) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar + p) : "memory");
}
__device__ __forceinline__ void phase_wait(const unsigned int *bar, int p, unsigned int G) {
    while (ld_acquire(bar + p) < G) { __nanosleep(20); }
}
__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

constexpr int TL_CW = 16;                          // math warps
constexpr int TL_CONSUMERS = TL_CW * 32;           // 512
constexpr int TL_THREADS = TL_CONSUMERS + 64;      // + copy warp + finishing warp
constexpr int TL_TILE = 1152;                      // bytes per Q4_0 tile (16 rows x 4 blocks)
// per weight type: tile bytes, tiles per math warp and ring slot, offset of the scale pairs inside a tile.
// Q8_0 tile (2176 B): 4 planes of 512 B (rows g bytes 0-15 | rows g bytes 16-31 | rows g+8 bytes 0-15 | rows g+8 bytes 16-31, 16 B per
// lane each: conflict-free LDS.128), then 32 x (fp16 d(row g), fp16 d(row g+8)).  An int8 XOR 0x80 in the low byte of a half is
// the fp16 subnormal (q + 128) x 2^-24: 0.75 ALU ops per weight, the +128 goes into the same per-block correction as Q4_0's 8.
template <int TYPE> struct TileCfg;
template <> struct TileCfg<NL_Q4_0> { static constexpr int TILE = 1152, TPW = 2, D_OFF = 1024; };
template <> struct TileCfg<NL_Q8_0> { static constexpr int TILE = 2176, TPW = 1, D_OFF = 2048; };
inline int tile_bytes(int type) { return type == NL_Q8_0 ? TileCfg<NL_Q8_0>::TILE : TileCfg<NL_Q4_0>::TILE; }
constexpr int TL_TS = 2 * TL_CW;                   // tiles per ring slot: two per math warp (a 31-tile slot that relieves the math warp
                                                   // next to the finishing warp was measured 15 % slower: slots stop lining up with the
                                                   // 32-tile row groups of 4096-column matrices, so fragments are reloaded every slot)
constexpr int TL_SLOT_BYTES = TL_TS * TL_TILE;     // 36,864

constexpr int TL_SLOTS = 4;
constexpr int TL_MAX_NBG = 96;                     // block groups per row: cols <= 12,288
// Shared-memory image of the phase input: per 32-element block a hi and a lo column of 16 words.  The B-fragment loads of the streaming
// loop read 8 columns at once (one lane each, 16 bytes): with 64-byte columns four lanes share a bank group (a 4-way conflict on the
// busiest shared-memory instruction of the loop); 80-byte columns spread the eight lanes over all 32 banks.
#ifndef NL_TL_XCOL
#define NL_TL_XCOL 80
#endif
constexpr int TL_XCOL = NL_TL_XCOL;                // bytes between the columns of the shared-memory fragment image (64 = dense)
constexpr int TL_XBG = 8 * TL_XCOL;                // ... between its block groups
constexpr int TL_XFRAG_BYTES = TL_MAX_NBG * TL_XBG > 57344 ? TL_MAX_NBG * TL_XBG : 57344;   // the attention phase reuses the buffer and
                                                   // needs 56 KB for 96 cached K/V rows
constexpr int TL_MAX_ITEMS = 3;                    // 8-float items per math thread in the prologue
constexpr int TL_CORR_BYTES = TL_MAX_NBG * 4 * 16;  // per 32-element block: {correction, 2^k / S} of its low and of its high 16 elements
constexpr size_t TL_DYN_SMEM = (size_t)TL_SLOTS * TL_SLOT_BYTES + TL_XFRAG_BYTES + TL_CORR_BYTES;
// Fragment image of a phase input in global memory ("producer-side fragments"): the phase that PRODUCES a vector also publishes it in
// the consuming GEMV's final form -- per block group (128 elements) the 512 bytes of fp16 hi | lo B fragments exactly as they sit in
// shared memory, then per 16-element half block one 16-byte record {correction, 2^k / S, sum of squares, 0}.  The consumer's prologue
// is load -> check for the sentinel -> st.shared instead of a 150-instruction conversion on every one of the 148 SMs.
constexpr int TL_IMG_BG = 512 + 4 * 2 * 16;        // image bytes per block group
inline size_t tile_img_bytes(int cols) { return (size_t)((cols / 32 + 3) / 4) * TL_IMG_BG; }

enum { TEPI_STORE = 0, TEPI_RESID = 1, TEPI_SWIGLU = 2 };

struct TilePhase {
    int kind;                 // PH_GEMV / PH_ATTN
    int layer;                // PH_ATTN
    const uint8_t *tiles;     // tiled matrix: n_rg row groups x nbg tiles
    int n_rg, nbg;
    unsigned int nbg_magic;   // ceil(2^32 / nbg): tile index -> row group without a divide
    int unit_rg;              // row groups per distribution unit (2 for gate/up: gate group, then the up group of the same rows)
    int units;                // n_rg / unit_rg: what the CTAs share out (must stay below 2^32 / grid^2, see band_of)
    int cols;                 // input length (multiple of 32)
    int rows;                 // valid output rows (<= 16 * n_rg / unit_rg for SWIGLU, <= 16 * n_rg otherwise)
    int epi;                  // TEPI_*
    const float *x;           // input vector [cols], fp32 (PH_ATTN: this layer's q | k | v vector); polled for the sentinel when in_poll
    const float *norm_w;      // non-null: input is RMSNorm(x; norm_w), go/quant.go:597-607
    const float *bias;        // optional [rows]
    float *out;               // output vector (PH_ATTN: the attention output); stored with st_poll when out_poll
    const float *resid;       // TEPI_RESID: the vector the product is added to (polled when resid_poll)
    int in_poll, out_poll, resid_poll;   // TileArgs::poll: the vector lives in the single-use arena (see "polled activations" in nl_tile.cu)
    const uint8_t *in_img;    // non-null: the input arrives as a polled fragment image (x / norm_w are then only used for their length)
    uint8_t *out_img;         // non-null: publish the outputs as the fragment image of the NEXT GEMV's input (PH_ATTN: the o-projection's)
    const float *out_nw;      // norm weights of that next GEMV (its input is RMSNorm(out; out_nw)), or null
    // ---- tensor parallel (TileArgs::tp > 1) ----
    int exch_out;             // row-split matrix (O / down): the product is this rank's PARTIAL; it is stored into slot `rank` of
                              // parity `par` of every peer's exchange area instead of being added to the residual
    int in_exch;              // input = prev + sum over ranks of the exchange area's parity `par` (fixed rank order: identical everywhere);
                              // CTA 0 also stores it to `next`, the residual of the following exchange
    int par;
    int cross;                // this phase's outputs travel to peers: its barrier counts every CTA of every rank
    int wait_cross;           // `cross` of the previous phase (whose barrier this phase waits on)
    const float *prev;
    float *next;
    // polled exchange (TileArgs::poll with tp > 1): exch_out stores go to window offset exch_off + (rank * dim + row) * 4 of EVERY rank
    // (this token parity's arena); an in_exch consumer polls prev (when prev_poll), the tp partial vectors at `parts` and leaves the sum
    // in `next` (polled by the exchange after it)
    unsigned long long exch_off;
    const float *parts;
    int prev_poll;
};

struct TileArgs {
    const TilePhase *phases;
    int n_phases;
    unsigned int *bar;        // [n_phases] grid-barrier counters, zeroed before every launch
    const unsigned int *epoch;  // launch counter behind the flags of the split-attention partials
    // tensor parallel: one process per GPU, peers' windows mapped through CUDA IPC (nl_tp.cuh); offsets are the same in every window
    int tp, rank, dim, lvocab;
    TpPeers peers;
    unsigned long long ar_off, bar_off, lg_off, amax_off;   // exchange area [2][tp][dim] f32 | barrier counters | full logits | [tp][grid] argmax pairs
    float2 *amax;             // optional [grid]: per CTA (maximum, index as int bits) of the last phase's outputs (device-side greedy)
    const int *lg_want;       // tensor parallel: nonzero = the caller reads the logits, every rank's shard goes to every window (else only
                              // the argmax pairs cross NVLink: the greedy / bench loops)
    int poll;                 // 1: activations are single-use polled vectors, the kernel has no grid barrier
    int slim;                 // 1: single GPU, polled, every GEMV input is an fp32 vector (no fragment images, no exchanges); 2: polled tensor
                              // parallel -- launch_tiled may take the kernel instantiation without the branches these runs never reach
    MegaAttn at;
    float eps;
    unsigned int g_magic;     // ceil(2^32 / grid), filled in by launch_tiled
    int poll_ns;              // back-off between two looks at a polled phase input (0 = look again at once)
    int att_chunk;            // attention: positions per split while the splits last (<= 96 = one pass)
    int att_hpi;              // attention: q heads per item; 0 = as few as still give every item its own CTA, >= group = the whole GQA group
    int inflight;             // ring copies requested but not yet landed, 1..TL_SLOTS
    int l2pf;                 // slots beyond the ring that the copy warp keeps requested in L2 (0 = no prefetch)
    int dbg;                  // forensics (NL_TILE_DBG; results are garbage): 1 = slots are handed over without copying (what the math
                              // warps and the phase boundaries cost on their own), 2 = every slot is copied from the band's first two
                              // slots (L2-resident source: the L2-fed rate), 3 = 1 + the finishing warp skips the partial sums of slots
                              // that do not complete a row group (-DNL_TL_DBG_SKIPMATH=1 drops the tile products too)
    unsigned long long *trace;  // optional: [cta][phase][8] globaltimer stamps
    unsigned long long *trace2; // optional: [cta][phase][16] clock64 stamps (tools/trace_fine.py)
};

// planar (qs, d) -> tiles: row group R of the source lands at tile row group rg_off + R * rg_stride (gate/up interleave: stride 2,
// offsets 0 / 1; q,k,v concatenation: stride 1, running offsets).  Rows / blocks beyond the matrix become zero blocks (d = 0).
int launch_tile_repack(int type, const uint8_t *qs, const __half *d, int rows, int nb, uint8_t *tiles, int nbg, int rg_off, int rg_stride, cudaStream_t st);
int launch_tiled(int type, const TileArgs &a, int grid, cudaStream_t st);
inline int tile_inflight() {
    const char *e = getenv("NL_TILE_INFLIGHT");
    int k = e ? atoi(e) : TL_SLOTS;
    return k < 1 ? 1 : (k > TL_SLOTS ? TL_SLOTS : k);
}
inline int tile_env_int(const char *name, int dflt, int lo, int hi) {
    const char *e = getenv(name);
    const int k = e ? atoi(e) : dflt;
    return k < lo ? lo : (k > hi ? hi : k);
}
inline unsigned int tile_magic(int nbg) { return nbg <= 1 ? 0u : (unsigned int)(((1ull << 32) + (unsigned)nbg - 1) / (unsigned)nbg); }

}  // namespace nl
