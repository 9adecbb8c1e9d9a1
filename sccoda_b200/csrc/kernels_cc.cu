// The case/control specialisation of the HMC kernel (fp32, one warp per chain, chain state in registers).
#include "hmc_cc.cuh"

namespace sccoda {
cudaError_t launch_hmc_cc(const KArgs& a, int grid, int block, size_t smem, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(hmc_cc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    hmc_cc_kernel<<<grid, block, smem, st>>>(a);
    return cudaGetLastError();
}
}  // namespace sccoda
