// fp64 instantiation of the sampler kernels, element path (parity build).
#include "kernels_impl.cuh"

namespace sccoda {
cudaError_t launch_kernel_f64(int kind, int W, const KArgs& a, int grid, int block, size_t smem, cudaStream_t st) {
    return launch_w<double, false>(kind, W, a, grid, block, smem, st);
}
}  // namespace sccoda
