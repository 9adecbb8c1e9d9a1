// The case/control specialisations of the HMC and NUTS kernels (fp32, one warp per chain, lane-per-cell-type state).
#include "hmc_cc.cuh"

namespace sccoda {
cudaError_t launch_hmc_cc(const KArgs& a, int grid, int block, size_t smem, cudaStream_t st) {
    // under-filled device: the builds with a larger register budget (hmc_cc.cuh), as long as every CTA of the launch is
    // still resident at once (3 CTAs of 128 threads per SM without a cap, 4 with 128 registers)
    const long long warps = (long long)grid * (block / 32);
    auto* fn = a.wide > 16 ? hmc_cc_wide_kernel : a.wide > 0 ? hmc_cc_wide16_kernel
                          : (warps <= 148LL * 12 ? hmc_cc_lat_kernel : (warps <= 148LL * 16 ? hmc_cc_lat16_kernel : hmc_cc_kernel));
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
