// Internal launch interface between capi.cu and kernels.cu / summary.cu.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

#include "model.cuh"

namespace sccoda {

enum { KIND_LOGP = 0, KIND_HMC = 1, KIND_NUTS = 2 };

struct KArgs {
    Dims dm;
    int B, C, Umax;
    int method, num_results, num_burnin, num_adapt, num_leapfrog, max_depth;
    int step_begin, step_end, store_draws;
    double step_size, target_accept;
    uint64_t seed;
    uint32_t chain_id0;
    const void* y;
    const void* ycst;
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

int nvec_for(int kind, int max_depth);
size_t smem_bytes_for(const KArgs& a, int kind, int W, int chains_per_cta, int dtype);
cudaError_t launch_kernel(int kind, int dtype, int W, int chains_per_cta, const KArgs& a, cudaStream_t st);

// K4 (summary.cu)
cudaError_t launch_summary(int dtype, int B, int C, int N, int K, int D, const int* ref, int num_results, int skip,
                           const void* draws, const void* stats, double* summary, cudaStream_t st);

}  // namespace sccoda
