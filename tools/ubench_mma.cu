// ubench_mma.cu — what the legacy tensor path (mma.sync.m16n8k16 f16 -> f32, SASS HMMA.16816.F32) costs on sm_100a: the latency of a
// dependent accumulator chain and the issue rate with 1..8 independent chains per warp, for 1 / 4 warps per SM sub-partition.  These
// two numbers decide how the decode kernel's streaming loop has to be shaped (nl_tile.cu, tile_dot: NL_TL_ACC).
//     nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/ubench_mma tools/ubench_mma.cu && gpurun_out/ubench_mma
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int CH>
__global__ void k_mma(long long *out, float *sink, int iters) {
    float c[CH][4];
#pragma unroll
    for (int k = 0; k < CH; k++) for (int j = 0; j < 4; j++) c[k][j] = 0.f;
    uint32_t a = threadIdx.x * 0x00010001u & 0x000F000Fu, b = 0x3C003C00u;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < CH; k++) mma(c[k], a, a, a, a, b, b);
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < CH; k++) s += c[k][0] + c[k][1] + c[k][2] + c[k][3];
    if (s == 12345.678f) sink[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

// the nibble extraction of tile_dot: 5 ALU ops per word (LOP3 / SHF), independent words
__global__ void k_alu(long long *out, uint32_t *sink, int iters) {
    uint32_t w[8], acc = 0;
    for (int i = 0; i < 8; i++) w[i] = threadIdx.x * 2654435761u + i;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint32_t a8 = w[i] >> 8;
            acc += (w[i] & 0x000F000Fu) ^ (w[i] & 0x00F000F0u) ^ (a8 & 0x000F000Fu) ^ (a8 & 0x00F000F0u);
            w[i] = w[i] * 3 + acc;
        }
    }
    const long long t1 = clock64();
    if (acc == 0x12345678u) sink[0] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

__global__ void k_shfl(long long *out, float *sink, int iters) {
    float v = threadIdx.x;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; i++) v += __shfl_xor_sync(0xffffffffu, v, 1);
    const long long t1 = clock64();
    if (v == 12345.678f) sink[0] = v;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

template <int CH> static void run(int threads, long long *d_out, float *d_sink) {
    const int iters = 4096;
    k_mma<CH><<<1, threads>>>(d_out, d_sink, iters);
    k_mma<CH><<<1, threads>>>(d_out, d_sink, iters);
    long long cyc = 0;
    cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost);
    const int warps_per_sp = (threads / 32 + 3) / 4;
    printf("{\"op\": \"HMMA.16816.F32\", \"chains_per_warp\": %d, \"warps\": %d, \"cycles_per_mma_per_warp\": %.2f, \"cycles_per_mma_per_subpartition\": %.2f}\n", CH, threads / 32,
           (double)cyc / iters / CH, (double)cyc / iters / CH / warps_per_sp);
}

int main() {
    long long *d_out; float *d_sink;
    cudaMalloc(&d_out, 8); cudaMalloc(&d_sink, 64);
    for (int threads : {32, 128, 512}) {
        run<1>(threads, d_out, d_sink); run<2>(threads, d_out, d_sink); run<4>(threads, d_out, d_sink); run<8>(threads, d_out, d_sink);
    }
    long long cyc = 0;
    for (int threads : {32, 128, 512}) {
        k_alu<<<1, threads>>>(d_out, (uint32_t *)d_sink, 4096);
        cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost);
        printf("{\"op\": \"nibble extraction (8 words: 8 SHF + 32 LOP3 + 8 IMAD + 32 IADD/LOP)\", \"warps\": %d, \"cycles_per_8_words_per_warp\": %.2f}\n", threads / 32, (double)cyc / 4096);
    }
    k_shfl<<<1, 32>>>(d_out, d_sink, 4096);
    cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost);
    printf("{\"op\": \"SHFL.BFLY + FADD dependent\", \"cycles\": %.2f}\n", (double)cyc / 4096);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
