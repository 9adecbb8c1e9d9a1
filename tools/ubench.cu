// Issue-rate micro-benchmarks for the pipes the sampler kernel lives on (tools, not product).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench tools/ubench.cu
// One CTA of 1024 threads per SM (8 warps per scheduler); rates are warp-instructions per clock per scheduler
// (SMSP), from clock64() inside the kernel.
#include <cuda_runtime.h>

#include <cstdio>

constexpr int ITERS = 2048;

template <int MODE>
__global__ void __launch_bounds__(1024) k(float* out, long long* cyc, float a, float b) {
    float x0 = threadIdx.x * 1e-3f + 1.f, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f,
          x6 = x0 + 6.f, x7 = x0 + 7.f;
    unsigned long long p0, p1, p2, p3, pa, pb;
    asm("mov.b64 %0, {%1,%2};" : "=l"(p0) : "f"(x0), "f"(x1));
    asm("mov.b64 %0, {%1,%2};" : "=l"(p1) : "f"(x2), "f"(x3));
    asm("mov.b64 %0, {%1,%2};" : "=l"(p2) : "f"(x4), "f"(x5));
    asm("mov.b64 %0, {%1,%2};" : "=l"(p3) : "f"(x6), "f"(x7));
    asm("mov.b64 %0, {%1,%2};" : "=l"(pa) : "f"(a), "f"(a));
    asm("mov.b64 %0, {%1,%2};" : "=l"(pb) : "f"(b), "f"(b));
    __shared__ float sm[2048];
    sm[threadIdx.x] = x0;
    sm[threadIdx.x + 1024] = x1;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < ITERS; ++i) {
        if (MODE == 0) {  // 8 FFMA
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        } else if (MODE == 1) {  // 4 FFMA2 (= 8 FMAs)
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p0) : "l"(pa), "l"(pb));
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p1) : "l"(pa), "l"(pb));
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p2) : "l"(pa), "l"(pb));
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p3) : "l"(pa), "l"(pb));
        } else if (MODE == 2) {  // 4 MUFU.RCP
            asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x0));
            asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x1));
            asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x2));
            asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x3));
        } else if (MODE == 3) {  // 4 MUFU.LG2
            asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(x0));
            asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(x1));
            asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(x2));
            asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(x3));
        } else if (MODE == 4) {  // 4 MUFU.EX2
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x0));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x1));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x2));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x3));
        } else if (MODE == 5) {  // 1 MUFU + 4 FFMA2 + 2 FFMA  (does the MUFU hide under FMA issue?)
            asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x6));
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p0) : "l"(pa), "l"(pb));
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p1) : "l"(pa), "l"(pb));
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p2) : "l"(pa), "l"(pb));
            x7 = fmaf(x7, a, b);
            x5 = fmaf(x5, a, b);
        } else if (MODE == 6) {  // 4 LDS.32 (conflict-free) + 4 FADD consuming them
            x0 += sm[(threadIdx.x + i) & 1023];
            x1 += sm[1024 + ((threadIdx.x + i) & 1023)];
            x2 += sm[(threadIdx.x + 2 * i) & 1023];
            x3 += sm[1024 + ((threadIdx.x + 3 * i) & 1023)];
        } else if (MODE == 7) {  // 4 FFMA + 4 IADD-like (alu pipe) : dual-pipe issue
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(*(unsigned*)&x4) : "r"(i), "r"(0x55u));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(*(unsigned*)&x5) : "r"(i), "r"(0x33u));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(*(unsigned*)&x6) : "r"(i), "r"(0x0fu));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(*(unsigned*)&x7) : "r"(i), "r"(0x77u));
        } else if (MODE == 8) {  // 2 FFMA2 + 2 MUFU : packed math next to the special-function pipe
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p0) : "l"(pa), "l"(pb));
            asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x4));
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p1) : "l"(pa), "l"(pb));
            asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(x5));
        }
    }
    const long long t1 = clock64();
    float r0, r1;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(r0), "=f"(r1) : "l"(p0));
    float s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7 + r0 + r1;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(r0), "=f"(r1) : "l"(p1));
    s += r0 + r1;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(r0), "=f"(r1) : "l"(p2));
    s += r0 + r1;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(r0), "=f"(r1) : "l"(p3));
    s += r0 + r1;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int instr_per_iter, float* out, long long* cyc, int sms) {
    for (int rep = 0; rep < 2; ++rep) k<MODE><<<sms, 1024>>>(out, cyc, 1.0000001f, 1e-7f);
    cudaDeviceSynchronize();
    long long h[1024];
    cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < sms; ++i) avg += (double)h[i];
    avg /= sms;
    // 32 warps per SM, 8 per scheduler
    const double per_smsp = 8.0 * ITERS * instr_per_iter / avg;
    printf("%-42s %7.3f warp-instr/clk/SMSP  (%d instr/iter, %.0f cycles)\n", name, per_smsp, instr_per_iter, avg);
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* out;
    long long* cyc;
    cudaMalloc(&out, sizeof(float) * sms * 1024);
    cudaMalloc(&cyc, sizeof(long long) * sms);
    run<0>("FFMA (3-reg)", 8, out, cyc, sms);
    run<1>("FFMA2 (f32x2)", 4, out, cyc, sms);
    run<2>("MUFU.RCP", 4, out, cyc, sms);
    run<3>("MUFU.LG2", 4, out, cyc, sms);
    run<4>("MUFU.EX2", 4, out, cyc, sms);
    run<5>("1 MUFU + 3 FFMA2 + 2 FFMA", 6, out, cyc, sms);
    run<6>("4 LDS.32 + 4 FADD (+index math)", 8, out, cyc, sms);
    run<7>("4 FFMA + 4 LOP3", 8, out, cyc, sms);
    run<8>("2 FFMA2 + 2 MUFU", 4, out, cyc, sms);
    printf("cuda status: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
