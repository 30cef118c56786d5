// per-token persistent decode kernel, NL_Q8_0 weights (see nl_mega.cuh)
#include "nl_mega.cuh"
namespace nl {
int launch_mega_q8_0(const MegaArgs &a, int grid, size_t smem, cudaStream_t st) { return launch_mega_typed<NL_Q8_0>(a, grid, smem, st); }
}  // namespace nl
