// The case/control specialisations of the HMC and NUTS kernels (fp32, one warp per chain, lane-per-cell-type state).
#include "hmc_cc.cuh"

namespace sccoda {
cudaError_t launch_hmc_cc(const KArgs& a, int grid, int block, size_t smem, cudaStream_t st) {
    // under-filled device (at most 16 warps per SM could ever be resident): the uncapped build (hmc_cc.cuh)
    const bool lat = (long long)grid * (block / 32) <= 148LL * 16;
    auto* fn = a.wide > 0 ? hmc_cc_wide_kernel : (lat ? hmc_cc_lat_kernel : hmc_cc_kernel);
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    fn<<<grid, block, smem, st>>>(a);
    return cudaGetLastError();
}
cudaError_t launch_logp_cc(const KArgs& a, int grid, int block, size_t smem, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(logp_grad_cc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    logp_grad_cc_kernel<<<grid, block, smem, st>>>(a);
    return cudaGetLastError();
}
cudaError_t launch_nuts_cc(const KArgs& a, int grid, int block, size_t smem, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(nuts_cc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    nuts_cc_kernel<<<grid, block, smem, st>>>(a);
    return cudaGetLastError();
}
}  // namespace sccoda
