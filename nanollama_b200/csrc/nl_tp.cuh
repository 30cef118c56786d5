// nl_tp.cuh — tensor-parallel exchange over NVLink peer memory (one process per GPU, CUDA IPC windows).
//
// The Go engine is single-process; splitting its Forward (go/model.go:490-620) Megatron-style over n GPUs needs a sum of the
// o-projection and down-projection partials (2 per layer) and a gather of the vocab-split logits.  Messages are tiny (16 KB at
// dim 4096), i.e. pure latency, so instead of a ring the exchange is ONE-SHOT: every rank stores its partial straight into a
// slot of every peer's window (NVSwitch gives all peers full bandwidth), raises a flag there, waits for its own n flags and
// sums the n slots in rank order (deterministic, identical on every rank), fused with the residual add.
#pragma once
#include "nl_common.cuh"

namespace nl {

constexpr int TP_MAX = 8;

// Layout of one rank's window (all offsets in bytes from the window base; identical on every rank)
struct TpLayout {
    size_t ar_data;   // float [2 parities][tp][dim]
    size_t ar_flag;   // uint32 [2][TP_MAX]   (padded to 128 B per parity)
    size_t lg_data;   // float [vocab]          full logits, every rank writes its shard into every window
    size_t lg_flag;   // uint32 [TP_MAX]
    size_t tile_bar;  // uint32 [1024]          phase-barrier counters of the persistent tiled kernel (never reset)
    size_t tile_amax; // float2 [TP_MAX][256]   per-CTA argmax pairs of every rank's LM-head shard
    size_t arena[2];  // polled activation arenas of the persistent tiled kernel, one per token parity (TpArena); peers store into them
    size_t pf_flag;   // uint32 [2][TP_MAX]     one-pass prefill: "partial ready" / "rows reduced" epochs of every rank
    size_t pf_part;   // float [rows][dim]      one-pass prefill: this rank's partial of a row-split GEMM (peers READ their row slice)
    size_t pf_x;      // float [rows][dim]      one-pass prefill: the residual stream (the owner of a row slice WRITES it into every window)
    size_t total;
};
// One parity of the tensor-parallel polled arena (nl_tile.cu, "polled activations" across ranks).  Everything a token's phases hand to
// each other lives here, is written exactly once per token and starts out as sentinels: local vectors (q|k|v, fragment images of the
// attention / SwiGLU outputs, the residual stream after each exchange) and the slots the PEERS store their partials of the row-split
// products into ([tp][dim] per exchange), plus every rank's per-CTA argmax pairs.  Two parities: a rank refills the one the NEXT token
// will use while the current token runs -- no peer writes it before every rank has finished the current token.
struct TpArena {
    size_t qkv, ao_img, part_o, xres, hb_img, part_d, xout, per_layer, amax, total;   // byte offsets (per layer / of the tail) and sizes
};
inline TpArena tp_arena(int tp, int dim, int nqkv, size_t img_q, size_t img_f, int n_layers) {
    TpArena a;
    size_t o = 0;
    a.qkv = o; o += (size_t)nqkv * 4;
    a.ao_img = o; o += img_q;
    a.part_o = o; o += (size_t)tp * dim * 4;
    a.xres = o; o += (size_t)dim * 4;
    a.hb_img = o; o += img_f;
    a.part_d = o; o += (size_t)tp * dim * 4;
    a.xout = o; o += (size_t)dim * 4;
    a.per_layer = (o + 255) / 256 * 256;
    a.amax = a.per_layer * n_layers;
    a.total = (a.amax + (size_t)TP_MAX * 256 * 8 + 255) / 256 * 256;
    return a;
}
inline TpLayout tp_layout(int tp, int dim, int vocab, size_t arena_bytes = 0, int pf_rows = 0) {
    TpLayout L;
    size_t o = 0;
    L.ar_data = o; o += (size_t)2 * tp * dim * 4; o = (o + 255) / 256 * 256;
    L.ar_flag = o; o += 2 * 128; o = (o + 255) / 256 * 256;
    L.lg_data = o; o += (size_t)vocab * 4; o = (o + 255) / 256 * 256;
    L.lg_flag = o; o += 128; o = (o + 255) / 256 * 256;
    L.tile_bar = o; o += 1024 * 4;
    L.tile_amax = o; o += (size_t)TP_MAX * 256 * 8; o = (o + 255) / 256 * 256;
    L.arena[0] = o; o += arena_bytes;
    L.arena[1] = o; o += arena_bytes;
    L.pf_flag = o; o += 2 * 128; o = (o + 255) / 256 * 256;
    L.pf_part = o; o += (size_t)pf_rows * dim * 4; o = (o + 255) / 256 * 256;
    L.pf_x = o; o += (size_t)pf_rows * dim * 4;
    L.total = (o + 255) / 256 * 256;
    return L;
}

struct TpPeers { uint8_t *win[TP_MAX]; };   // win[r] = base of rank r's window as mapped into THIS process (win[rank] is local)

__device__ __forceinline__ void st_release_sys(unsigned int *p, unsigned int v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// x[i] += sum_r partial_r[i]   (X += WO·xb2 / X += WDown·hb with the matrices split along their input dimension)
// One CTA.  `epoch` lives in device memory and is advanced by the kernel itself so the launch can sit in a CUDA graph.
static __global__ void __launch_bounds__(1024) tp_allreduce_resid_kernel(const float *__restrict__ partial, float *__restrict__ x, int dim, TpPeers peers,
                                                                  TpLayout L, int rank, int tp, unsigned int *epoch) {
    const unsigned int e = *epoch + 1u;
    const int par = e & 1u;
    const int n4 = dim >> 2;
    // 1. my partial into slot `rank` of every window (remote stores ride NVLink)
    for (int r = 0; r < tp; r++) {
        float4 *dst = reinterpret_cast<float4 *>(peers.win[r] + L.ar_data) + ((size_t)par * tp + rank) * n4;
        const float4 *src = reinterpret_cast<const float4 *>(partial);
        for (int i = threadIdx.x; i < n4; i += blockDim.x) dst[i] = src[i];
    }
    __threadfence_system();
    __syncthreads();
    // 2. raise my flag in every window, 3. wait for everybody's flag in mine
    if (threadIdx.x < tp) {
        st_release_sys(reinterpret_cast<unsigned int *>(peers.win[threadIdx.x] + L.ar_flag) + par * 32 + rank, e);
        const unsigned int *mine = reinterpret_cast<const unsigned int *>(peers.win[rank] + L.ar_flag) + par * 32 + threadIdx.x;
        while ((int)(ld_acquire_sys(mine) - e) < 0) { }
    }
    __syncthreads();
    // 4. fixed-order sum + residual
    const float4 *slots = reinterpret_cast<const float4 *>(peers.win[rank] + L.ar_data) + (size_t)par * tp * n4;
    for (int i = threadIdx.x; i < n4; i += blockDim.x) {
        float4 acc = __ldcg(slots + i);
        for (int r = 1; r < tp; r++) {
            const float4 v = __ldcg(slots + (size_t)r * n4 + i);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        float4 xv = reinterpret_cast<float4 *>(x)[i];
        xv.x += acc.x; xv.y += acc.y; xv.z += acc.z; xv.w += acc.w;
        reinterpret_cast<float4 *>(x)[i] = xv;
    }
    __syncthreads();
    if (threadIdx.x == 0) *epoch = e;
}

// ---- one-pass prefill under tensor parallelism: X[T, dim] += sum over ranks of the partial products of a row-split GEMM (o / down).
// Bandwidth-sized (T * dim * 4 bytes), so it is a reduce-scatter + all-gather over peer memory instead of the one-shot exchange of the
// decode path: rank r owns the rows [r * T / tp, (r + 1) * T / tp); it LOADS that slice of every rank's partial over NVLink, adds them
// in rank order to the residual rows it holds (fixed order, computed once: identical everywhere) and STORES the new rows into every
// rank's residual buffer.  Two flags per rank bracket it: "my partial is complete" before anybody reads it, "my rows are everywhere"
// before anybody uses X or overwrites its partial (tp_rows_done_kernel, a separate launch: the grid has to have finished storing).
static __global__ void __launch_bounds__(256) tp_reduce_rows_kernel(int T, int dim, TpPeers peers, TpLayout L, int rank, int tp, const unsigned int *epoch) {
    const unsigned int e = *epoch + 1u;
    if (threadIdx.x < tp) {
        if (blockIdx.x == 0) st_release_sys(reinterpret_cast<unsigned int *>(peers.win[threadIdx.x] + L.pf_flag) + rank, e);   // (kernel boundary: the GEMM's stores are complete)
        const unsigned int *mine = reinterpret_cast<const unsigned int *>(peers.win[rank] + L.pf_flag) + threadIdx.x;
        while ((int)(ld_acquire_sys(mine) - e) < 0) { }
    }
    __syncthreads();
    const int r0 = (int)((long long)T * rank / tp), r1 = (int)((long long)T * (rank + 1) / tp);
    const size_t n4 = (size_t)(r1 - r0) * dim / 4, base4 = (size_t)r0 * dim / 4;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 acc = reinterpret_cast<const float4 *>(peers.win[rank] + L.pf_x)[base4 + i];
        for (int r = 0; r < tp; r++) {
            const float4 v = reinterpret_cast<const float4 *>(peers.win[r] + L.pf_part)[base4 + i];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        for (int r = 0; r < tp; r++) reinterpret_cast<float4 *>(peers.win[r] + L.pf_x)[base4 + i] = acc;
    }
}
static __global__ void tp_rows_done_kernel(TpPeers peers, TpLayout L, int rank, int tp, unsigned int *epoch) {
    const unsigned int e = *epoch + 1u;
    __threadfence_system();
    if (threadIdx.x < tp) {
        st_release_sys(reinterpret_cast<unsigned int *>(peers.win[threadIdx.x] + L.pf_flag) + 32 + rank, e);
        const unsigned int *mine = reinterpret_cast<const unsigned int *>(peers.win[rank] + L.pf_flag) + 32 + threadIdx.x;
        while ((int)(ld_acquire_sys(mine) - e) < 0) { }
    }
    __syncthreads();
    if (threadIdx.x == 0) *epoch = e;
}

// every rank's logits shard [lvocab] -> the full logits vector in every window
static __global__ void __launch_bounds__(1024) tp_allgather_logits_kernel(const float *__restrict__ local, int lvocab, TpPeers peers, TpLayout L, int rank, int tp,
                                                                   unsigned int *epoch) {
    const unsigned int e = *epoch + 1u;
    const int n4 = lvocab >> 2;
    for (int r = 0; r < tp; r++) {
        float4 *dst = reinterpret_cast<float4 *>(peers.win[r] + L.lg_data) + (size_t)rank * n4;
        const float4 *src = reinterpret_cast<const float4 *>(local);
        for (int i = threadIdx.x; i < n4; i += blockDim.x) dst[i] = src[i];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < tp) {
        st_release_sys(reinterpret_cast<unsigned int *>(peers.win[threadIdx.x] + L.lg_flag) + rank, e);
        const unsigned int *mine = reinterpret_cast<const unsigned int *>(peers.win[rank] + L.lg_flag) + threadIdx.x;
        while ((int)(ld_acquire_sys(mine) - e) < 0) { }
    }
    __syncthreads();
    if (threadIdx.x == 0) *epoch = e;
}

}  // namespace nl
