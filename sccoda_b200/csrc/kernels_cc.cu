// The case/control specialisations of the HMC and NUTS kernels (fp32, one warp per chain, lane-per-cell-type state).
#include "hmc_cc.cuh"

namespace sccoda {
cudaError_t launch_hmc_cc(const KArgs& a, int grid, int block, size_t smem, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(hmc_cc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    hmc_cc_kernel<<<grid, block, smem, st>>>(a);
    return cudaGetLastError();
}
cudaError_t launch_nuts_cc(const KArgs& a, int grid, int block, size_t smem, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(nuts_cc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    nuts_cc_kernel<<<grid, block, smem, st>>>(a);
    return cudaGetLastError();
}
}  // namespace sccoda
