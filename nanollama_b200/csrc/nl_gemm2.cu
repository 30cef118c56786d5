// second-generation tcgen05 GEMM instantiations (see nl_gemm2.cuh)
#include "nl_gemm2.cuh"
namespace nl {
int launch_gemm2_q4_0(const Gemm2Args &g, cudaStream_t st) { return launch_gemm2_typed<NL_Q4_0>(g, st); }
int launch_gemm2_q8_0(const Gemm2Args &g, cudaStream_t st) { return launch_gemm2_typed<NL_Q8_0>(g, st); }
// F16 rows are 64 bytes per K step: the raw ring next to two 128-token tiles would be two steps deep, so the wide shape takes one tile per CTA
int launch_gemm2_f16(const Gemm2Args &g, cudaStream_t st) { return g.T > 128 ? launch_gemm2_t<NL_F16, 1>(g, st) : launch_gemm2_t<NL_F16, 0>(g, st); }
}  // namespace nl
