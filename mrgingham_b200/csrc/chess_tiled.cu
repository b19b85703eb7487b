// K1 tiled variant -- placeholder that forwards to the simple kernel until the tiled kernel lands.
#include "kernels.cuh"
namespace mrgb200
{
cudaError_t launch_chess_sparse_tiled(const FrameSet& fs, cand_t* cand, uint32_t* counts,
                                      int cand_capacity, cudaStream_t stream)
{
    return launch_chess_sparse_simple(fs, cand, counts, cand_capacity, stream);
}
}
