// The multi-group HMC and NUTS kernels (fp32, one warp per chain, lane-per-cell-type state), one instantiation per covariate count.
#include "hmc_mg.cuh"

namespace sccoda {
template <int D>
static cudaError_t launch_hmc_mg_d(const KArgs& a, int grid, int block, size_t smem, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(hmc_mg_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    hmc_mg_kernel<D><<<grid, block, smem, st>>>(a);
    return cudaGetLastError();
}
cudaError_t launch_hmc_mg(const KArgs& a, int grid, int block, size_t smem, cudaStream_t st) {
    switch (a.dm.D) {
        case 1: return launch_hmc_mg_d<1>(a, grid, block, smem, st);
        case 2: return launch_hmc_mg_d<2>(a, grid, block, smem, st);
        case 3: return launch_hmc_mg_d<3>(a, grid, block, smem, st);
        case 4: return launch_hmc_mg_d<4>(a, grid, block, smem, st);
        default: return cudaErrorInvalidValue;
    }
}
template <int D>
static cudaError_t launch_nuts_mg_d(const KArgs& a, int grid, int block, size_t smem, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(nuts_mg_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    nuts_mg_kernel<D><<<grid, block, smem, st>>>(a);
    return cudaGetLastError();
}
cudaError_t launch_nuts_mg(const KArgs& a, int grid, int block, size_t smem, cudaStream_t st) {
    switch (a.dm.D) {
        case 1: return launch_nuts_mg_d<1>(a, grid, block, smem, st);
        case 2: return launch_nuts_mg_d<2>(a, grid, block, smem, st);
        case 3: return launch_nuts_mg_d<3>(a, grid, block, smem, st);
        case 4: return launch_nuts_mg_d<4>(a, grid, block, smem, st);
        default: return cudaErrorInvalidValue;
    }
}
}  // namespace sccoda
