// K4 — per-chain posterior summaries straight from the device-resident draws.
//
// Replaces the host-side python loops of get_y_hat (sccoda/model/scCODA_model.py:806-820: ind, b_raw, beta per
// draw) and CAResult.complete_beta_df (sccoda/util/result_classes.py:243-255: inclusion probability and
// non-zero mean per effect) plus the moments az.summary needs (mean / sd), so that a 1024-dataset sweep never
// has to move 19 GB of draws to the host.  HBM-bound: every draw is read exactly once, coalesced.
#include <cuda_runtime.h>

#include "../../include/sccoda_b200.h"
#include "kernels.h"
#include "special.cuh"

namespace sccoda {

// One CTA (8 warps) per chain.  Outputs ("tasks") come in three kinds — state coordinates, beta_dk, scalar
// statistics of the trace — and every warp works on ONE kind (no divergence): a warp takes 32 consecutive
// coordinates (or 32 effects) and one slice of the rows, walking the rows four at a time with independent
// accumulators (the loads of a warp are consecutive floats of one draw: coalesced); the statistics warp spreads the
// rows over its lanes.  Partial sums meet in shared memory and are combined in slice order (deterministic).
constexpr int kSummaryThreads = 256;
constexpr int kSummaryWarps = kSummaryThreads / 32;

struct SumAcc {
    double a, b, c, d;   // coordinate: (s1, s2, -, -); beta: (s1, s2, count, included sum); scalar: (v, -, -, -)
};

template <typename T>
__global__ void __launch_bounds__(kSummaryThreads) summary_kernel(int C, int K, int D, const int* __restrict__ refs, int S,
                                                                  int skip, const T* __restrict__ draws,
                                                                  const T* __restrict__ stats, double* __restrict__ summary) {
    extern __shared__ __align__(16) unsigned char summary_smem[];
    SumAcc* part = reinterpret_cast<SumAcc*>(summary_smem);   // [n_slices][n_tasks]
    const int chain = blockIdx.x;
    const int b = chain / C;
    const int K1 = K - 1, DK1 = D * K1, P = D + 2 * DK1 + K, DK = D * K;
    const int n_tasks = P + DK + SCCODA_NSUM_SCALAR;
    const int ncg = (P + 31) / 32, nbg = (DK + 31) / 32, n_units = ncg + nbg + 1;
    const int n_slices = n_units >= kSummaryWarps ? 1 : kSummaryWarps / n_units;
    const int ref = refs[b];
    const T* dr = draws + (size_t)chain * S * P;
    const T* st = stats + (size_t)chain * S * SCCODA_NSTAT;
    double* out = summary + (size_t)chain * (SCCODA_NSUM_PARAM * P + SCCODA_NSUM_BETA * DK + SCCODA_NSUM_SCALAR);
    const int n_used = S - skip;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    for (int wu = warp; wu < n_units * n_slices; wu += kSummaryWarps) {
        const int slice = wu / n_units, unit = wu - slice * n_units;
        const bool post = unit < ncg + nbg;
        // rows of this slice: posterior tasks walk [skip, S), the statistics of the whole trace walk [0, S)
        const int lo_all = post ? skip : 0;
        const int per = (S - lo_all + n_slices - 1) / n_slices;
        const int r0 = lo_all + slice * per, r1 = min(S, r0 + per);
        SumAcc* ps = part + (size_t)slice * n_tasks;
        if (unit < ncg) {
            // mean and M2 of one state coordinate (sums shifted by the first used draw, fp64)
            const int task = unit * 32 + lane;
            if (task < P) {
                const double x0 = n_used > 0 ? (double)dr[(size_t)skip * P + task] : 0.0;
                double s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
                int r = r0;
                for (; r + 3 < r1; r += 4) {
                    T v[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) v[j] = dr[(size_t)(r + j) * P + task];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const double w = (double)v[j] - x0;
                        s1[j] += w;
                        s2[j] = fma(w, w, s2[j]);
                    }
                }
                for (; r < r1; ++r) {
                    const double w = (double)dr[(size_t)r * P + task] - x0;
                    s1[0] += w;
                    s2[0] = fma(w, w, s2[0]);
                }
                ps[task] = SumAcc{(s1[0] + s1[1]) + (s1[2] + s1[3]), (s2[0] + s2[1]) + (s2[2] + s2[3]), 0.0, 0.0};
            }
        } else if (post) {
            // beta_dk = sigmoid(50 ind_raw) * sigma_d * b_offset, zero at the reference column
            const int e = (unit - ncg) * 32 + lane;
            if (e < DK) {
                const int d = e / K, k = e - d * K;
                SumAcc acc = {0.0, 0.0, 0.0, 0.0};
                if (k != ref) {
                    const int j = k - (k > ref);
                    const int i_sig = d, i_off = D + d * K1 + j, i_raw = D + DK1 + d * K1 + j;
#pragma unroll 4
                    for (int r = r0; r < r1; ++r) {
                        const T* row = dr + (size_t)r * P;
                        const T ind = sigmoid_stable(T(50) * row[i_raw]);
                        const double beta = (double)(ind * (row[i_sig] * row[i_off]));
                        if (fabs(beta) > 1e-3) {  // result_classes.py:249
                            acc.c += 1.0;
                            acc.d += beta;
                        }
                        acc.a += beta;
                        acc.b = fma(beta, beta, acc.b);
                    }
                }
                ps[P + e] = acc;
            }
        } else {
            // accepted steps, sum exp(min(log_accept_ratio, 0)), leapfrogs, diverging steps: rows over the lanes
            double v1 = 0.0, v2 = 0.0, v3 = 0.0, v4 = 0.0;
            for (int r = r0 + lane; r < r1; r += 32) {
                const T* row = st + (size_t)r * SCCODA_NSTAT;
                v1 += (double)row[SCCODA_ST_IS_ACCEPTED];
                v2 += exp(fmin((double)row[SCCODA_ST_LOG_ACCEPT_RATIO], 0.0));
                v3 += (double)row[SCCODA_ST_LEAPFROGS_TAKEN];
                v4 += (double)row[SCCODA_ST_DIVERGING];
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                v1 += __shfl_xor_sync(0xffffffffu, v1, o);
                v2 += __shfl_xor_sync(0xffffffffu, v2, o);
                v3 += __shfl_xor_sync(0xffffffffu, v3, o);
                v4 += __shfl_xor_sync(0xffffffffu, v4, o);
            }
            if (lane < SCCODA_NSUM_SCALAR) {
                const double v = lane == 1 ? v1 : lane == 2 ? v2 : lane == 3 ? v3 : lane == 4 ? v4 : 0.0;
                ps[P + DK + lane] = SumAcc{v, 0.0, 0.0, 0.0};
            }
        }
    }
    __syncthreads();
    for (int task = threadIdx.x; task < n_tasks; task += kSummaryThreads) {
        SumAcc acc = part[task];
        for (int sl = 1; sl < n_slices; ++sl) {
            const SumAcc o = part[(size_t)sl * n_tasks + task];
            acc.a += o.a; acc.b += o.b; acc.c += o.c; acc.d += o.d;
        }
        if (task < P) {
            const double x0 = n_used > 0 ? (double)dr[(size_t)skip * P + task] : 0.0;
            const double mean = n_used > 0 ? acc.a / n_used : 0.0;
            out[task] = x0 + mean;
            out[P + task] = acc.b - acc.a * mean;
        } else if (task < P + DK) {
            const int e = task - P;
            const double mean = n_used > 0 ? acc.a / n_used : 0.0;
            double* ob = out + SCCODA_NSUM_PARAM * P;
            ob[e] = acc.c;
            ob[DK + e] = acc.d;
            ob[2 * DK + e] = mean;
            ob[3 * DK + e] = acc.b - acc.a * mean;
        } else {
            const int q = task - P - DK;
            double v = acc.a;
            if (q == 0) v = (double)n_used;
            else if (q == 5) v = S > 0 ? (double)st[(size_t)(S - 1) * SCCODA_NSTAT + SCCODA_ST_STEP_SIZE] : 0.0;
            out[SCCODA_NSUM_PARAM * P + SCCODA_NSUM_BETA * DK + q] = v;
        }
    }
}

template <typename T>
static cudaError_t launch_summary_t(int grid, size_t smem, cudaStream_t st, int C, int K, int D, const int* ref, int S, int skip,
                                    const void* draws, const void* stats, double* summary) {
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(summary_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    summary_kernel<T><<<grid, kSummaryThreads, smem, st>>>(C, K, D, ref, S, skip, static_cast<const T*>(draws),
                                                            static_cast<const T*>(stats), summary);
    return cudaGetLastError();
}

cudaError_t launch_summary(int dtype, int B, int C, int N, int K, int D, const int* ref, int num_results, int skip,
                           const void* draws, const void* stats, double* summary, cudaStream_t st) {
    (void)N;
    const int grid = B * C;
    const int P = D + 2 * D * (K - 1) + K, DK = D * K, n_tasks = P + DK + SCCODA_NSUM_SCALAR;
    const int n_units = (P + 31) / 32 + (DK + 31) / 32 + 1;
    const int n_slices = n_units >= kSummaryWarps ? 1 : kSummaryWarps / n_units;
    const size_t smem = sizeof(SumAcc) * (size_t)n_slices * n_tasks;   // partial sums [n_slices][n_tasks]
    if (dtype == SCCODA_F64) return launch_summary_t<double>(grid, smem, st, C, K, D, ref, num_results, skip, draws, stats, summary);
    return launch_summary_t<float>(grid, smem, st, C, K, D, ref, num_results, skip, draws, stats, summary);
}

}  // namespace sccoda
