// dequant-fused GEMV instantiations for NL_Q4_0 weights (see nl_kernels.cuh / nl_gemv.cuh)
#include "nl_gemv.cuh"
namespace nl {
int launch_gemv_q4_0(const GemvArgs &a, int batch, int epi, cudaStream_t st) { return launch_gemv_typed<NL_Q4_0>(a, batch, epi, st); }
}  // namespace nl
