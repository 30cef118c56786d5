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
    size_t total;
};
inline TpLayout tp_layout(int tp, int dim, int vocab) {
    TpLayout L;
    size_t o = 0;
    L.ar_data = o; o += (size_t)2 * tp * dim * 4; o = (o + 255) / 256 * 256;
    L.ar_flag = o; o += 2 * 128; o = (o + 255) / 256 * 256;
    L.lg_data = o; o += (size_t)vocab * 4; o = (o + 255) / 256 * 256;
    L.lg_flag = o; o += 128; o = (o + 255) / 256 * 256;
    L.tile_bar = o; o += 1024 * 4;
    L.tile_amax = o; o += (size_t)TP_MAX * 256 * 8;
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
