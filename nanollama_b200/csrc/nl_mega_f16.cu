// per-token persistent decode kernel, NL_F16 weights (see nl_mega.cuh)
#include "nl_mega.cuh"
namespace nl {
int launch_mega_f16(const MegaArgs &a, int grid, size_t smem, cudaStream_t st) { return launch_mega_typed<NL_F16>(a, grid, smem, st); }
}  // namespace nl
