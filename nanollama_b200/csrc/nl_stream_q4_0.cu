// streaming GEMV instantiations for NL_Q4_0 weights (see nl_stream.cuh)
#include "nl_stream.cuh"
namespace nl {
int launch_stream_q4_0(const StreamArgs &a, int NM, int RPT, int act, int grid, size_t smem, cudaStream_t st, bool pdl) {
    return launch_stream_typed<NL_Q4_0>(a, NM, RPT, act, grid, smem, st, pdl);
}
}  // namespace nl
