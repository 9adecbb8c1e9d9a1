// fp64 instantiation of the sampler kernels (parity build: follows the CPU oracle's trajectories).
#include "kernels_impl.cuh"

namespace sccoda {
cudaError_t launch_kernel_f64(int kind, int W, const KArgs& a, int grid, int block, size_t smem, cudaStream_t st) {
    return launch_w<double>(kind, W, a, grid, block, smem, st);
}
}  // namespace sccoda
