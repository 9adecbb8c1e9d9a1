// The multi-group HMC and NUTS kernels (fp32, one warp per chain, lane-per-cell-type state), one instantiation per
// covariate count D and per number of cell types a lane owns (NS = 1: K <= 31, NS = 2: K <= 62).
#include "hmc_mg.cuh"

namespace sccoda {
template <int D, int NS>
static cudaError_t launch_hmc_mg_d(const KArgs& a, int grid, int block, size_t smem, cudaStream_t st) {
    if (a.wide > 0) {
        cudaError_t e = cudaFuncSetAttribute(hmc_mg_wide_kernel<D, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        hmc_mg_wide_kernel<D, NS><<<grid, block, smem, st>>>(a);
        return cudaGetLastError();
    }
    cudaError_t e = cudaFuncSetAttribute(hmc_mg_kernel<D, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    hmc_mg_kernel<D, NS><<<grid, block, smem, st>>>(a);
    return cudaGetLastError();
}
template <int D, int NS>
static cudaError_t launch_nuts_mg_d(const KArgs& a, int grid, int block, size_t smem, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(nuts_mg_kernel<D, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    nuts_mg_kernel<D, NS><<<grid, block, smem, st>>>(a);
    return cudaGetLastError();
}
template <int D, int NS>
static cudaError_t launch_logp_mg_d(const KArgs& a, int grid, int block, size_t smem, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(logp_grad_mg_kernel<D, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    logp_grad_mg_kernel<D, NS><<<grid, block, smem, st>>>(a);
    return cudaGetLastError();
}
cudaError_t launch_logp_mg(const KArgs& a, int grid, int block, size_t smem, cudaStream_t st) {
    if (a.dm.K <= 31) {
        switch (a.dm.D) {
            case 1: return launch_logp_mg_d<1, 1>(a, grid, block, smem, st);
            case 2: return launch_logp_mg_d<2, 1>(a, grid, block, smem, st);
            case 3: return launch_logp_mg_d<3, 1>(a, grid, block, smem, st);
            case 4: return launch_logp_mg_d<4, 1>(a, grid, block, smem, st);
        }
    } else {
        switch (a.dm.D) {
            case 1: return launch_logp_mg_d<1, 2>(a, grid, block, smem, st);
            case 2: return launch_logp_mg_d<2, 2>(a, grid, block, smem, st);
            case 3: return launch_logp_mg_d<3, 2>(a, grid, block, smem, st);
        }
    }
    return cudaErrorInvalidValue;
}
cudaError_t launch_hmc_mg(const KArgs& a, int grid, int block, size_t smem, cudaStream_t st) {
    if (a.dm.K <= 31) {
        switch (a.dm.D) {
            case 1: return launch_hmc_mg_d<1, 1>(a, grid, block, smem, st);
            case 2: return launch_hmc_mg_d<2, 1>(a, grid, block, smem, st);
            case 3: return launch_hmc_mg_d<3, 1>(a, grid, block, smem, st);
            case 4: return launch_hmc_mg_d<4, 1>(a, grid, block, smem, st);
        }
    } else {
        switch (a.dm.D) {
            case 1: return launch_hmc_mg_d<1, 2>(a, grid, block, smem, st);
            case 2: return launch_hmc_mg_d<2, 2>(a, grid, block, smem, st);
            case 3: return launch_hmc_mg_d<3, 2>(a, grid, block, smem, st);
        }
    }
    return cudaErrorInvalidValue;
}
cudaError_t launch_nuts_mg(const KArgs& a, int grid, int block, size_t smem, cudaStream_t st) {
    if (a.dm.K <= 31) {
        switch (a.dm.D) {
            case 1: return launch_nuts_mg_d<1, 1>(a, grid, block, smem, st);
            case 2: return launch_nuts_mg_d<2, 1>(a, grid, block, smem, st);
            case 3: return launch_nuts_mg_d<3, 1>(a, grid, block, smem, st);
            case 4: return launch_nuts_mg_d<4, 1>(a, grid, block, smem, st);
        }
    } else {
        switch (a.dm.D) {
            case 1: return launch_nuts_mg_d<1, 2>(a, grid, block, smem, st);
            case 2: return launch_nuts_mg_d<2, 2>(a, grid, block, smem, st);
            case 3: return launch_nuts_mg_d<3, 2>(a, grid, block, smem, st);
        }
    }
    return cudaErrorInvalidValue;
}
}  // namespace sccoda
