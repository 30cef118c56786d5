// second-generation tcgen05 GEMM instantiations (see nl_gemm2.cuh)
#include "nl_gemm2.cuh"
namespace nl {
int launch_gemm2_q4_0(const Gemm2Args &g, cudaStream_t st) { return launch_gemm2_typed<NL_Q4_0>(g, st); }
int launch_gemm2_q8_0(const Gemm2Args &g, cudaStream_t st) { return launch_gemm2_typed<NL_Q8_0>(g, st); }
int launch_gemm2_f16(const Gemm2Args &g, cudaStream_t st) { return g.T > 128 ? -3 : launch_gemm2_t<NL_F16, 0>(g, st); }
}  // namespace nl
