// K5 — posterior-predictive quantities per draw, straight from the device-resident draws.
//
// Replaces the per-draw python loops of get_y_hat (sccoda/model/scCODA_model.py:806-827): for every stored draw
//   ind = sigmoid(50 ind_raw), beta = ind * sigma_d * b_offset (zero at the reference cell type)      (:806-820)
//   concentration[n,k] = exp(alpha_k + x_n . beta_k)                                                  (:821-822)
//   prediction[n,k]    = n_total_n * concentration[n,k] / sum_k concentration[n,k]  (DirMult mean)    (:824-827)
// so that the [S,N,K] arrays are produced on demand (a range of draws at a time) instead of being materialised on
// the host for the whole chain.  HBM-write-bound: one warp per draw, every output row written once, coalesced.
#include <cuda_runtime.h>

#include "../../include/sccoda_b200.h"
#include "kernels.h"
#include "special.cuh"

namespace sccoda {

constexpr int kPredictWarps = 8;

__device__ __forceinline__ float p_exp(float v) { return expf(v); }
__device__ __forceinline__ double p_exp(double v) { return exp(v); }

// floats of scratch one warp needs: theta[P] | beta[D*K] | c[Umax*K] | S[Umax]
__host__ __device__ inline size_t predict_warp_elems(int K, int D, int Umax) {
    const size_t P = (size_t)D + 2 * (size_t)D * (K - 1) + K;
    return (P + (size_t)D * K + (size_t)Umax * K + Umax + 3) & ~size_t(3);
}

template <typename T>
__global__ void __launch_bounds__(kPredictWarps * 32) predict_kernel(int C, int N, int K, int D, int Umax, const int* __restrict__ refs,
                                                                     const T* __restrict__ xu_all, const int* __restrict__ grp_all,
                                                                     const T* __restrict__ ntot_all, const int* __restrict__ order_all, int S,
                                                                     int lo, int hi,
                                                                     int draws_per_cta, int warps, const T* __restrict__ draws,
                                                                     T* __restrict__ conc, T* __restrict__ pred) {
    extern __shared__ __align__(16) unsigned char predict_smem[];
    const int n_tiles = (hi - lo + draws_per_cta - 1) / draws_per_cta;
    const int chain = blockIdx.x / n_tiles, tile = blockIdx.x - chain * n_tiles;
    const int b = chain / C;
    const int K1 = K - 1, DK1 = D * K1, P = D + 2 * DK1 + K, NK = N * K;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // per CTA: covariate values of the groups; group, total and output position of every (packed) row
    T* xu = reinterpret_cast<T*>(predict_smem);
    T* ntot = xu + Umax * D;
    int* grp = reinterpret_cast<int*>(ntot + N);
    int* dst = grp + N;
    T* scratch = reinterpret_cast<T*>(predict_smem + (((size_t)(Umax * D + N) * sizeof(T) + (size_t)N * 8 + 15) & ~size_t(15)));
    for (int i = threadIdx.x; i < Umax * D; i += blockDim.x) xu[i] = xu_all[(size_t)b * Umax * D + i];
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        ntot[i] = ntot_all[(size_t)b * N + i];
        grp[i] = grp_all[(size_t)b * N + i];
        dst[i] = order_all ? order_all[(size_t)b * N + i] : i;     // the packed rows are sorted by covariate group
    }
    __syncthreads();
    if (warp >= warps) return;
    const int ref = refs[b];
    T* th = scratch + (size_t)warp * predict_warp_elems(K, D, Umax);
    T* beta = th + P;
    T* cu = beta + D * K;
    T* Su = cu + Umax * K;
    const int s_end = min(hi, lo + (tile + 1) * draws_per_cta);
    for (int s = lo + tile * draws_per_cta + warp; s < s_end; s += warps) {
        const T* row = draws + ((size_t)chain * S + s) * P;
        for (int i = lane; i < P; i += 32) th[i] = row[i];
        __syncwarp();
        for (int e = lane; e < D * K; e += 32) {
            const int d = e / K, k = e - d * K;
            T v = T(0);
            if (k != ref) {
                const int j = k - (k > ref);
                v = sigmoid_stable(T(50) * th[D + DK1 + d * K1 + j]) * (th[d] * th[D + d * K1 + j]);
            }
            beta[e] = v;
        }
        __syncwarp();
        for (int e = lane; e < Umax * K; e += 32) {
            const int u = e / K, k = e - u * K;
            T lin = th[D + 2 * DK1 + k];
            for (int d = 0; d < D; ++d) lin = fma(xu[u * D + d], beta[d * K + k], lin);
            cu[e] = p_exp(lin);
        }
        __syncwarp();
        for (int u = 0; u < Umax; ++u) {
            T part = T(0);
            for (int k = lane; k < K; k += 32) part += cu[u * K + k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
            if (lane == 0) Su[u] = part;
        }
        __syncwarp();
        const size_t o = ((size_t)chain * (hi - lo) + (s - lo)) * NK;
        for (int idx = lane; idx < NK; idx += 32) {
            const int n = idx / K, k = idx - n * K, u = grp[n];
            const T c = cu[u * K + k];
            const size_t w = o + (size_t)dst[n] * K + k;
            if (conc) conc[w] = c;
            if (pred) pred[w] = ntot[n] * c / Su[u];
        }
        __syncwarp();
    }
}

template <typename T>
static cudaError_t launch_predict_t(int B, int C, int N, int K, int D, int Umax, const int* ref, const void* xu, const int* grp,
                                    const void* ntot, const int* order, int S, int lo, int hi, const void* draws, void* conc, void* pred,
                                    cudaStream_t st) {
    int dev = 0, max_smem = 0, sms = 148;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const size_t head = (((size_t)(Umax * D + N) * sizeof(T) + (size_t)N * 8 + 15) & ~size_t(15));
    const size_t per_warp = predict_warp_elems(K, D, Umax) * sizeof(T);
    int warps = kPredictWarps;
    while (warps > 1 && head + warps * per_warp > (size_t)max_smem) warps >>= 1;
    const size_t smem = head + warps * per_warp;
    if (smem > (size_t)max_smem) return cudaErrorInvalidConfiguration;
    // enough CTAs to fill the device a few times over, at least one draw per warp
    const int chains = B * C, n = hi - lo;
    int tiles = (sms * 8 + chains - 1) / chains;
    if (tiles > (n + warps - 1) / warps) tiles = (n + warps - 1) / warps;
    if (tiles < 1) tiles = 1;
    const int draws_per_cta = (n + tiles - 1) / tiles;
    tiles = (n + draws_per_cta - 1) / draws_per_cta;
    if (smem > 48 * 1024) {
        e = cudaFuncSetAttribute(predict_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    predict_kernel<T><<<chains * tiles, kPredictWarps * 32, smem, st>>>(
        C, N, K, D, Umax, ref, static_cast<const T*>(xu), grp, static_cast<const T*>(ntot), order, S, lo, hi, draws_per_cta, warps,
        static_cast<const T*>(draws), static_cast<T*>(conc), static_cast<T*>(pred));
    return cudaGetLastError();
}

cudaError_t launch_predict(int dtype, int B, int C, int N, int K, int D, int Umax, const int* ref, const void* xu, const int* grp,
                           const void* ntot, const int* order, int num_results, int lo, int hi, const void* draws, void* conc,
                           void* pred, cudaStream_t st) {
    if (dtype == SCCODA_F64)
        return launch_predict_t<double>(B, C, N, K, D, Umax, ref, xu, grp, ntot, order, num_results, lo, hi, draws, conc, pred, st);
    return launch_predict_t<float>(B, C, N, K, D, Umax, ref, xu, grp, ntot, order, num_results, lo, hi, draws, conc, pred, st);
}

}  // namespace sccoda
