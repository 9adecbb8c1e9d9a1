// fp32 instantiation of the sampler kernels, grouped (pair / item) path: the production build for designs with few covariate groups.
#include "kernels_impl.cuh"

namespace sccoda {
cudaError_t launch_kernel_f32g(int kind, int W, const KArgs& a, int grid, int block, size_t smem, cudaStream_t st) {
    return launch_w<float, true>(kind, W, a, grid, block, smem, st);
}
}  // namespace sccoda
