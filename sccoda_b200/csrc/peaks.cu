// Roofline denominators measured on the device the bench runs on: FP32 FFMA issue peak and MUFU (SFU) peak.
// MEASURED_PEAKS.json carries only HBM and bf16-GEMM numbers; this path is FP32/SFU-bound, so the
// bench measures its own denominators with these two micro-kernels and reports fractions "of measured".
#include <cuda_runtime.h>

#include "../../include/sccoda_b200.h"

namespace {

constexpr int ITERS = 4096;

__global__ void __launch_bounds__(256) ffma_peak_kernel(float* out, float a, float b) {
    float x0 = threadIdx.x * 1e-3f, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f,
          x6 = x0 + 6.f, x7 = x0 + 7.f;
#pragma unroll 16
    for (int i = 0; i < ITERS; ++i) {
        x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
        x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

__global__ void __launch_bounds__(256) mufu_peak_kernel(float* out, float a) {
    float x0 = threadIdx.x * 1e-3f + a, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f;
#pragma unroll 16
    for (int i = 0; i < ITERS; ++i) {
        // four independent chains; no rcp(rcp(x)) pairs, which ptxas folds away
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x0));
        asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(x1));
        asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(x2));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x3));
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3;
}

}  // namespace

extern "C" int sccoda_measure_peaks(float* fp32_tflops, float* mufu_tops, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -100;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = sms * 8, block = 256;
    float* buf = nullptr;
    if (cudaMalloc(&buf, sizeof(float) * grid * block) != cudaSuccess) return -42;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best_f = 1e30f, best_m = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {
        float ms = 0.f;
        cudaEventRecord(e0, st);
        ffma_peak_kernel<<<grid, block, 0, st>>>(buf, 1.0000001f, 1e-7f);
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best_f) best_f = ms;
        cudaEventRecord(e0, st);
        mufu_peak_kernel<<<grid, block, 0, st>>>(buf, 0.5f);
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best_m) best_m = ms;
    }
    const double threads = (double)grid * block;
    if (fp32_tflops) *fp32_tflops = (float)(threads * ITERS * 8 * 2 / (best_f * 1e-3) / 1e12);
    if (mufu_tops) *mufu_tops = (float)(threads * ITERS * 4 / (best_m * 1e-3) / 1e12);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(buf);
    return cudaGetLastError() == cudaSuccess ? 0 : -100;
}
