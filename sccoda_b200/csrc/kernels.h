// Internal launch interface between capi.cu and the kernel translation units.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

namespace sccoda {

enum {
    KIND_LOGP = 0, KIND_HMC = 1, KIND_NUTS = 2,
    KIND_HMC_CC = 3, KIND_NUTS_CC = 4,   // case/control specialisations (hmc_cc.cuh)
    KIND_HMC_MG = 5, KIND_NUTS_MG = 6    // up to 32 covariate groups, D <= 4, K <= 62 on the same lane-per-cell-type layout (hmc_mg.cuh)
};

}  // namespace sccoda

#include "model.cuh"

namespace sccoda {

struct KArgs {
    Dims dm;
    int B, C, Umax;
    int el_smem;  // 1: stage the element list into shared memory (fp32 build, when it fits); 0: read it from global
    int el_count; // entries staged when el_smem: N*K, or N*K - nbig_min with the grouped layout (counts < 6 only)
    int grouped;  // 1: pair / item gradient path (grouped.cuh); 0: element path (model.cuh)
    int NImax, NOmax, NGmax, NF, NT, NOT;
    const int* NI;
    const void* iy;
    const uint16_t* ipair;
    const uint16_t* ipos;
    const uint8_t* inreal;
    const void* pcoef;
    const int* NG;
    const void* oy;
    const uint16_t* opair;
    const uint16_t* opos;
    const uint32_t* odesc;
    int method, num_results, num_burnin, num_adapt, num_leapfrog, max_depth;
    int step_begin, step_end, store_draws;
    int generic_only;  // 1: never take a specialised kernel (tests compare the specialisations with the generic path)
    int wide;          // > 0: the wide HMC launch (hmc_cc_wide_kernel, hmc_mg_wide_kernel): warps (= chains) of the fullest CTA
    int wide_grid;     //      and its number of CTAs (one per SM)
    double step_size, target_accept;
    uint64_t seed;
    uint32_t chain_id0;
    const void* ey;
    const void* eycst;
    const uint32_t* eidx;
    const int* colperm;
    const int* nbig;
    const void* ntot;
    const void* xu;
    const int* grp;
    const int* rstart;
    const int* U;
    const int* ref;
    const double* lconst;
    const void* init;
    void* draws;
    void* stats;
    double* carry;
};

// ------------------------------------------------------------------------------------------------
// shared-memory plan (identical on host and device)
// ------------------------------------------------------------------------------------------------
__host__ __device__ inline size_t rnd16(size_t bytes) { return (bytes + 15) & ~size_t(15); }

struct SmemPlan {
    size_t ntot, nrc, xu, grp, rstart, colperm, el, eyc, pcoef, iy, ipair, ipos, inreal, stage_red, data_total;   // per CTA
    size_t red, cp, rowterm, G, Gu, beta, ind, gk, vec, chain_total;                  // per chain (relative)
};

// counts per item of the grouped layout (SCCODA_ITEM_R of grouped.cuh / include/sccoda_b200.h)
constexpr int kItemR = 5;

// nimax > 0 selects the grouped (pair / item) layout: NP = Umax*(K+1) pair records instead of (c, psi) pairs, nf
// result slots per pair (+ one spare) instead of the per-element G, no row terms and no column-sum scratch.
__host__ __device__ inline SmemPlan plan_smem(int N, int K, int D, int Umax, int el_count, int nimax, int nf, int nvec,
                                              int W, size_t ts) {
    SmemPlan p;
    const size_t NK = (size_t)N * K, UK = (size_t)Umax * K, P = (size_t)D + 2 * (size_t)D * (K - 1) + K;
    const size_t NP = (size_t)Umax * (K + 1);
    const bool grouped = nimax > 0;
    size_t o = 0;
    p.ntot = o; o += rnd16((size_t)N * ts);
    p.nrc = o; o += rnd16((size_t)N * ts);                       // per-row constants of the fp32 row terms (value pass)
    p.xu = o; o += rnd16((size_t)Umax * D * ts);
    p.grp = o; o += rnd16((size_t)N * 4);
    p.rstart = o; o += rnd16((size_t)(Umax + 1) * 4);
    p.colperm = o; o += rnd16((size_t)K * 4);
    p.el = o; o += rnd16((size_t)el_count * 2 * ts);             // el_count = 0: the element list stays in global memory
    p.eyc = o; o += rnd16((size_t)el_count * ts);
    p.pcoef = o; o += grouped ? rnd16(NP * 6 * ts) : 0;
    p.iy = o; o += grouped ? rnd16((size_t)kItemR * nimax * ts) : 0;
    p.ipair = o; o += grouped ? rnd16((size_t)nimax * 2) : 0;
    p.ipos = o; o += grouped ? rnd16((size_t)nimax * 2) : 0;
    p.inreal = o; o += grouped ? rnd16((size_t)nimax) : 0;
    p.stage_red = o; o += 32 * 8;                                // double[32]: one slot per warp for the staging sum
    p.data_total = o;
    o = 0;
    p.red = o; o += rnd16((size_t)2 * W * 8);
    p.cp = o; o += grouped ? rnd16(NP * 4 * ts) : rnd16(UK * 2 * ts);
    p.rowterm = o; o += grouped ? 0 : rnd16((size_t)N * ts);
    p.G = o; o += grouped ? rnd16((NP * nf + 1) * ts) : rnd16(NK * ts);
    p.Gu = o; o += rnd16((size_t)Umax * ts);
    p.beta = o; o += rnd16((size_t)D * K * ts);
    p.ind = o; o += rnd16((size_t)D * (K - 1) * ts + ts);
    p.gk = o; o += grouped ? 0 : rnd16((size_t)(D + 1) * K * ts);
    p.vec = o; o += rnd16((size_t)nvec * rnd16(P * ts));
    p.chain_total = o;
    return p;
}

// Shared-memory plan of the case/control HMC kernel (hmc_cc.cuh): lane-major arrays at compile-time offsets first.
struct CcPlan {
    size_t cf, yv, ov, om, rows, el, bounds, red, fill, data_total;   // per CTA
    size_t pc, vec, slots, chain_total;        // per chain (relative)
    int NTC;                                   // item slots per pair: the packed items plus room for the shifted odd counts
};
// wide launch: datasets a CTA of `warps` consecutive chains (C per dataset) can span at most
__host__ __device__ inline int wide_datasets(int warps, int C) { return (warps + C - 2) / C + 1; }
// n_slots: float4-per-lane vectors of the NUTS tree (0 for HMC)
__host__ __device__ inline CcPlan plan_cc(int N, int K, int NT, int NOT, int n_slots) {
    CcPlan p;
    // odd counts (the 0.5 pseudocount, non-integer data) are staged as counts y + 6 in the item slots of their pair
    // (hmc_cc.cuh: cc_stage); the padding of the last packed item takes them, at worst ceil(NOT / R) more items
    p.NTC = NT + (NOT + kItemR - 1) / kItemR;
    size_t o = 0;
    p.cf = o; o += 6 * 32 * 8;                                   // f32x2[6][32] rational coefficients (group 0, group 1)
    p.yv = o; o += (size_t)p.NTC * kItemR * 32 * 8;              // f32x2[NTC][R][32] item counts
    p.ov = o; o += (size_t)NOT * 32 * 8;                         // f32x2[NOT][32] distinct odd counts of a pair ...
    p.om = o; o += (size_t)NOT * 32 * 8;                         // f32x2[NOT][32] ... and how often each occurs
    p.rows = o; o += (size_t)N * 16;                             // (row total, host-like constant, group, -)
    p.el = o; o += (size_t)N * K * 16;                           // (count, host constant, concentration slot, -)
    p.bounds = o; o += 16;                                       // int[3]: per-dataset loop bounds (items, odd counts, big elements)
    p.red = o; o += 32 * 8;                                      // double[32]: staging reduction, one slot per warp
    p.fill = o; o += 64 * 4;                                     // int[32][2]: data counts in the item slots of every pair
    p.data_total = o;
    p.pc = 0;                                                    // float[64] concentrations, slot u*32 + column
    p.vec = 256;                                                 // float[96] flat state vector
    p.slots = 256 + 384;                                         // float4[n_slots][32] stored vectors of the NUTS tree
    p.chain_total = 256 + 384 + (size_t)n_slots * 512;
    return p;
}

// Shared-memory plan of the multi-group kernels (hmc_mg.cuh): the arrays of CcPlan, one block per (PAIR of groups,
// column slot) — UP = pairs of groups of the fullest dataset, NS = cell types per lane (1: K <= 31, 2: K <= 62) —
// plus the covariate values and row counts of the groups.
struct MgPlan {
    size_t cf, yv, ov, om, rows, el, xg, cnt, bounds, red, fill, data_total;   // per CTA
    size_t pc, vec, slots, chain_total;                               // per chain (relative)
    int NTC;                                                          // item slots per pair incl. room for the shifted odd counts (CcPlan)
};
constexpr int kMgMaxGroups = 32, kMgMaxD = 4, kMgMaxK = 62, kMgMaxD2 = 3 /* largest D with two cell types per lane */;
// resident CTAs of 128 threads per SM the multi-group kernels are compiled for (register caps); the wide HMC launch is
// ONE CTA of as many warps (4 x mg_hmc_ctas)
__host__ __device__ constexpr int mg_hmc_ctas(int D, int NS) { return NS == 1 ? (D == 1 ? 6 : (D == 2 ? 5 : (D == 3 ? 4 : 3))) : (D == 1 ? 4 : (D == 2 ? 3 : 2)); }
__host__ __device__ constexpr int mg_nuts_ctas(int D, int NS) { return NS == 1 ? (D == 1 ? 4 : (D <= 2 ? 3 : 2)) : (D == 1 ? 3 : (D == 2 ? 2 : 1)); }
// float4 per lane of one stored state vector: NS (alpha + D b_offset + D ind_raw) + D sigma
__host__ __device__ constexpr int mg_vec_quads(int D, int NS) { return (NS * (1 + 2 * D) + D + 3) / 4; }
// n_slots: stored vectors of the NUTS tree (0 for HMC)
__host__ __device__ inline MgPlan plan_mg(int N, int K, int D, int UP, int NS, int NT, int NOT, int n_slots) {
    MgPlan p;
    const size_t Q = (size_t)UP * NS;
    p.NTC = NT + (NOT + kItemR - 1) / kItemR;
    size_t o = 0;
    p.cf = o; o += Q * 6 * 32 * 8;                               // f32x2[UP][NS][6][32] rational coefficients
    p.yv = o; o += Q * p.NTC * kItemR * 32 * 8;                  // f32x2[UP][NS][NTC][R][32] item counts
    p.ov = o; o += Q * NOT * 32 * 8;                             // f32x2[UP][NS][NOT][32] distinct odd counts of a pair ...
    p.om = o; o += Q * NOT * 32 * 8;                             // f32x2[UP][NS][NOT][32] ... and how often each occurs
    p.rows = o; o += (size_t)N * 16;                             // (row total, host-like constant, slot of S_u, -)
    p.el = o; o += (size_t)N * K * 16;                           // (count, host constant, concentration slot, 1/count)
    p.xg = o; o += rnd16((size_t)UP * D * 8);                    // f32x2[UP][D] covariate values of the groups
    p.cnt = o; o += rnd16((size_t)UP * 2 * 4);                   // float[2 UP] rows per group
    p.bounds = o; o += rnd16((size_t)(2 * Q + 1) * 4);           // int: items / odd counts per (group pair, slot), big elements
    p.red = o; o += 32 * 8;                                      // double[32]: staging reduction, one slot per warp
    p.fill = o; o += Q * 64 * 4;                                 // int[UP][NS][32][2]: data counts in the item slots of every pair
    p.data_total = o;
    const size_t P = (size_t)D + 2 * (size_t)D * (K - 1) + K;
    p.pc = 0;                                                    // float[2 UP][32 NS] concentrations, slot u*32*NS + column
    p.vec = Q * 256;                                             // flat state vector
    p.slots = p.vec + rnd16(P * 4);                              // float4[n_slots][quads][32]
    p.chain_total = p.slots + (size_t)n_slots * mg_vec_quads(D, NS) * 512;
    return p;
}

int nvec_for(int kind, int max_depth);
// KIND_HMC -> KIND_HMC_CC when the problem has the case/control shape the specialised kernel is written for
int select_kind(const KArgs& a, int kind, int dtype, int W);
cudaError_t launch_hmc_cc(const KArgs& a, int grid, int block, size_t smem, cudaStream_t st);
cudaError_t launch_nuts_cc(const KArgs& a, int grid, int block, size_t smem, cudaStream_t st);
cudaError_t launch_hmc_mg(const KArgs& a, int grid, int block, size_t smem, cudaStream_t st);
cudaError_t launch_logp_cc(const KArgs& a, int grid, int block, size_t smem, cudaStream_t st);
cudaError_t launch_logp_mg(const KArgs& a, int grid, int block, size_t smem, cudaStream_t st);
// does the problem have the shape the case/control (cc) / multi-group (mg) lane-per-cell-type kernels are written for?
bool fits_cc(const KArgs& a, int dtype);
bool fits_mg(const KArgs& a, int dtype);
cudaError_t launch_nuts_mg(const KArgs& a, int grid, int block, size_t smem, cudaStream_t st);
size_t smem_bytes_for(const KArgs& a, int kind, int W, int chains_per_cta, int dtype);
cudaError_t launch_kernel(int kind, int dtype, int W, int chains_per_cta, const KArgs& a, cudaStream_t st);
cudaError_t launch_kernel_f32(int kind, int W, const KArgs& a, int grid, int block, size_t smem, cudaStream_t st);
cudaError_t launch_kernel_f64(int kind, int W, const KArgs& a, int grid, int block, size_t smem, cudaStream_t st);
cudaError_t launch_kernel_f32g(int kind, int W, const KArgs& a, int grid, int block, size_t smem, cudaStream_t st);
cudaError_t launch_kernel_f64g(int kind, int W, const KArgs& a, int grid, int block, size_t smem, cudaStream_t st);

// K4 (summary.cu)
cudaError_t launch_summary(int dtype, int B, int C, int N, int K, int D, const int* ref, int num_results, int skip,
                           const void* draws, const void* stats, double* summary, cudaStream_t st);

// K5 (predict.cu)
cudaError_t launch_predict(int dtype, int B, int C, int N, int K, int D, int Umax, const int* ref, const void* xu, const int* grp,
                           const void* ntot, const int* order, int num_results, int lo, int hi, const void* draws, void* conc,
                           void* pred, cudaStream_t st);

}  // namespace sccoda
