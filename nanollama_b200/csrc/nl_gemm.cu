// tcgen05 prefill GEMM instantiations (see nl_gemm.cuh)
#include "nl_gemm.cuh"
namespace nl {
int launch_gemm_q4_0(const GemmArgs &g, cudaStream_t st) { return launch_gemm_typed<NL_Q4_0>(g, st); }
int launch_gemm_q8_0(const GemmArgs &g, cudaStream_t st) { return launch_gemm_typed<NL_Q8_0>(g, st); }
int launch_gemm_f16(const GemmArgs &g, cudaStream_t st) { return launch_gemm_typed<NL_F16>(g, st); }
}  // namespace nl
