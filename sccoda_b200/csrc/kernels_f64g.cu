// fp64 instantiation of the sampler kernels, grouped (pair / item) path (parity build).
#include "kernels_impl.cuh"

namespace sccoda {
cudaError_t launch_kernel_f64g(int kind, int W, const KArgs& a, int grid, int block, size_t smem, cudaStream_t st) {
    return launch_w<double, true>(kind, W, a, grid, block, smem, st);
}
}  // namespace sccoda
