// streaming GEMV instantiations for NL_F16 weights (see nl_stream.cuh)
#include "nl_stream.cuh"
namespace nl {
int launch_stream_f16(const StreamArgs &a, int NM, int RPT, int act, int grid, size_t smem, cudaStream_t st, bool pdl) {
    return launch_stream_typed<NL_F16>(a, NM, RPT, act, grid, smem, st, pdl);
}
}  // namespace nl
