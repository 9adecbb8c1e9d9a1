// Internal launch interface between capi.cu and the kernel translation units.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

namespace sccoda {

enum { KIND_LOGP = 0, KIND_HMC = 1, KIND_NUTS = 2 };

}  // namespace sccoda

#include "model.cuh"

namespace sccoda {

struct KArgs {
    Dims dm;
    int B, C, Umax;
    int el_smem;  // 1: stage the element list into shared memory (fp32 build, when it fits); 0: read it from global
    int method, num_results, num_burnin, num_adapt, num_leapfrog, max_depth;
    int step_begin, step_end, store_draws;
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
    size_t ntot, xu, grp, rstart, colperm, el, data_total;                        // per CTA
    size_t red, cp, rowterm, G, Gu, beta, ind, gk, vec, chain_total;       // per chain (relative)
};

__host__ __device__ inline SmemPlan plan_smem(int N, int K, int D, int Umax, int el_smem, int nvec, int W, size_t ts) {
    SmemPlan p;
    const size_t NK = (size_t)N * K, UK = (size_t)Umax * K, P = (size_t)D + 2 * (size_t)D * (K - 1) + K;
    size_t o = 0;
    p.ntot = o; o += rnd16((size_t)N * ts);
    p.xu = o; o += rnd16((size_t)Umax * D * ts);
    p.grp = o; o += rnd16((size_t)N * 4);
    p.rstart = o; o += rnd16((size_t)(Umax + 1) * 4);
    p.colperm = o; o += rnd16((size_t)K * 4);
    p.el = o; o += el_smem ? rnd16(NK * 2 * ts) : 0;
    p.data_total = o;
    o = 0;
    p.red = o; o += rnd16((size_t)2 * W * 8);
    p.cp = o; o += rnd16(UK * 2 * ts);
    p.rowterm = o; o += rnd16((size_t)N * ts);
    p.G = o; o += rnd16(NK * ts);
    p.Gu = o; o += rnd16((size_t)Umax * ts);
    p.beta = o; o += rnd16((size_t)D * K * ts);
    p.ind = o; o += rnd16((size_t)D * (K - 1) * ts + ts);
    p.gk = o; o += rnd16((size_t)(D + 1) * K * ts);
    p.vec = o; o += rnd16((size_t)nvec * rnd16(P * ts));
    p.chain_total = o;
    return p;
}

int nvec_for(int kind, int max_depth);
size_t smem_bytes_for(const KArgs& a, int kind, int W, int chains_per_cta, int dtype);
cudaError_t launch_kernel(int kind, int dtype, int W, int chains_per_cta, const KArgs& a, cudaStream_t st);
cudaError_t launch_kernel_f32(int kind, int W, const KArgs& a, int grid, int block, size_t smem, cudaStream_t st);
cudaError_t launch_kernel_f64(int kind, int W, const KArgs& a, int grid, int block, size_t smem, cudaStream_t st);

// K4 (summary.cu)
cudaError_t launch_summary(int dtype, int B, int C, int N, int K, int D, const int* ref, int num_results, int skip,
                           const void* draws, const void* stats, double* summary, cudaStream_t st);

}  // namespace sccoda
