// nl_gemv.cuh — launcher for the dequant-fused GEMV; instantiated once per weight type in nl_gemv_<type>.cu
#pragma once
#include "nl_gemv_kernel.cuh"

namespace nl {

// rows-per-CTA / threads-per-CTA heuristics: keep >= ~2 waves of CTAs on 148 SMs where the matrix allows it, and
// give every thread at least one 16-byte unit of the row.
struct GemvPlan { int R, threads; };
inline GemvPlan plan_gemv(int type, int max_rows, int cols, int epi) {
    const int ue = type == NL_F16 ? 8 : type == NL_F32 ? 4 : 16;
    const int units = cols / ue;
    GemvPlan p;
    p.threads = units <= 64 ? 64 : units <= 128 ? 128 : 256;
    p.R = (epi == EPI_SWIGLU) ? 4 : (max_rows >= 2048 ? 8 : 4);
    return p;
}

template <int TYPE, int R, int NB, int EPI>
static int launch_threads(const GemvArgs &a, int threads, int ctas, cudaStream_t st) {
    switch (threads) {
    case 64: gemv_kernel<TYPE, R, NB, 64, EPI><<<ctas, 64, 0, st>>>(a); break;
    case 128: gemv_kernel<TYPE, R, NB, 128, EPI><<<ctas, 128, 0, st>>>(a); break;
    default: gemv_kernel<TYPE, R, NB, 256, EPI><<<ctas, 256, 0, st>>>(a); break;
    }
    return 0;
}
template <int TYPE, int NB>
static int launch_nb(const GemvArgs &a, const GemvPlan &p, int epi, int ctas, cudaStream_t st) {
    if (epi == EPI_SWIGLU) return launch_threads<TYPE, 4, NB, EPI_SWIGLU>(a, p.threads, ctas, st);
    if (epi == EPI_RESID) return p.R == 8 ? launch_threads<TYPE, 8, NB, EPI_RESID>(a, p.threads, ctas, st)
                                          : launch_threads<TYPE, 4, NB, EPI_RESID>(a, p.threads, ctas, st);
    return p.R == 8 ? launch_threads<TYPE, 8, NB, EPI_STORE>(a, p.threads, ctas, st)
                    : launch_threads<TYPE, 4, NB, EPI_STORE>(a, p.threads, ctas, st);
}

// a: segments filled except cta_begin; batch in {1,2,4}
template <int TYPE>
int launch_gemv_typed(GemvArgs a, int batch, int epi, cudaStream_t st) {
    int max_rows = 0;
    for (int i = 0; i < a.nseg; i++) max_rows = a.seg[i].rows > max_rows ? a.seg[i].rows : max_rows;
    GemvPlan p = plan_gemv(TYPE, max_rows, a.cols, epi);
    int ctas = 0;
    for (int i = 0; i < a.nseg; i++) { a.seg[i].cta_begin = ctas; ctas += (a.seg[i].rows + p.R - 1) / p.R; }
    switch (batch) {
    case 1: return launch_nb<TYPE, 1>(a, p, epi, ctas, st);
    case 2: return launch_nb<TYPE, 2>(a, p, epi, ctas, st);
    case 4: return launch_nb<TYPE, 4>(a, p, epi, ctas, st);
    default: return -1;
    }
}

int launch_gemv_q4_0(const GemvArgs &a, int batch, int epi, cudaStream_t st);
int launch_gemv_q8_0(const GemvArgs &a, int batch, int epi, cudaStream_t st);
int launch_gemv_f16(const GemvArgs &a, int batch, int epi, cudaStream_t st);
int launch_gemv_f32(const GemvArgs &a, int batch, int epi, cudaStream_t st);

}  // namespace nl
