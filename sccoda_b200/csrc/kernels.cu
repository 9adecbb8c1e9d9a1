// Host-side dispatch shared by both dtype instantiations.
#include "kernels.h"

#include "../../include/sccoda_b200.h"

namespace sccoda {

int nvec_for(int kind, int max_depth) {
    if (kind == KIND_LOGP) return 2;
    if (kind == KIND_HMC) return 6;
    if (kind == KIND_HMC_CC || kind == KIND_NUTS_CC || kind == KIND_HMC_MG || kind == KIND_NUTS_MG) return 1;
    return 15 + 2 * (max_depth + 1);
}

bool fits_cc(const KArgs& a, int dtype) {
    return a.grouped && dtype == SCCODA_F32 && a.dm.D == 1 && a.Umax == 2 && a.dm.K <= 31;
}
bool fits_mg(const KArgs& a, int dtype) {
    return a.grouped && dtype == SCCODA_F32 && a.Umax <= kMgMaxGroups &&
           (a.dm.K <= 31 ? a.dm.D <= kMgMaxD : (a.dm.K <= kMgMaxK && a.dm.D <= kMgMaxD2));
}

int select_kind(const KArgs& a, int kind, int dtype, int W) {
    if ((kind == KIND_HMC || kind == KIND_NUTS) && !a.generic_only && W == 1 && fits_cc(a, dtype))
        return kind == KIND_HMC ? KIND_HMC_CC : KIND_NUTS_CC;
    if ((kind == KIND_HMC || kind == KIND_NUTS) && !a.generic_only && W == 1 && fits_mg(a, dtype))
        return kind == KIND_HMC ? KIND_HMC_MG : KIND_NUTS_MG;
    return kind;
}

size_t smem_bytes_for(const KArgs& a, int kind, int W, int chains_per_cta, int dtype) {
    const size_t ts = dtype == SCCODA_F64 ? 8 : 4;
    if (kind == KIND_HMC_CC || kind == KIND_NUTS_CC) {
        const CcPlan cp = plan_cc(a.dm.N, a.dm.K, a.NT, a.NOT, kind == KIND_NUTS_CC ? 11 + 2 * (a.max_depth + 1) : 0);
        if (kind == KIND_HMC_CC && a.wide > 0) return (size_t)wide_datasets(a.wide, a.C) * cp.data_total + (size_t)a.wide * cp.chain_total;
        return cp.data_total + (size_t)chains_per_cta * cp.chain_total;
    }
    if (kind == KIND_HMC_MG || kind == KIND_NUTS_MG) {
        const MgPlan mp = plan_mg(a.dm.N, a.dm.K, a.dm.D, (a.Umax + 1) / 2, a.dm.K <= 31 ? 1 : 2, a.NT, a.NOT,
                                  kind == KIND_NUTS_MG ? 11 + 2 * (a.max_depth + 1) : 0);
        if (kind == KIND_HMC_MG && a.wide > 0) return (size_t)wide_datasets(a.wide, a.C) * mp.data_total + (size_t)a.wide * mp.chain_total;
        return mp.data_total + (size_t)chains_per_cta * mp.chain_total;
    }
    const SmemPlan pl = plan_smem(a.dm.N, a.dm.K, a.dm.D, a.Umax, a.el_smem ? a.el_count : 0, a.grouped ? a.NImax : 0, a.NF,
                                  nvec_for(kind, a.max_depth), W, ts);
    return pl.data_total + (size_t)chains_per_cta * pl.chain_total;
}

cudaError_t launch_kernel(int kind, int dtype, int W, int chains_per_cta, const KArgs& a, cudaStream_t st) {
    const int ctas_per_ds = (a.C + chains_per_cta - 1) / chains_per_cta;
    const int grid = a.B * ctas_per_ds;
    const int block = chains_per_cta * 32 * W;
    const size_t smem = smem_bytes_for(a, kind, W, chains_per_cta, dtype);
    if (kind == KIND_HMC_CC && a.wide > 0) return launch_hmc_cc(a, a.wide_grid, a.wide * 32, smem, st);
    if (kind == KIND_HMC_CC) return launch_hmc_cc(a, grid, block, smem, st);
    if (kind == KIND_NUTS_CC) return launch_nuts_cc(a, grid, block, smem, st);
    if (kind == KIND_HMC_MG && a.wide > 0) return launch_hmc_mg(a, a.wide_grid, a.wide * 32, smem, st);
    if (kind == KIND_HMC_MG) return launch_hmc_mg(a, grid, block, smem, st);
    if (kind == KIND_NUTS_MG) return launch_nuts_mg(a, grid, block, smem, st);
    if (a.grouped) {
        if (dtype == SCCODA_F64) return launch_kernel_f64g(kind, W, a, grid, block, smem, st);
        return launch_kernel_f32g(kind, W, a, grid, block, smem, st);
    }
    if (dtype == SCCODA_F64) return launch_kernel_f64(kind, W, a, grid, block, smem, st);
    return launch_kernel_f32(kind, W, a, grid, block, smem, st);
}

}  // namespace sccoda
