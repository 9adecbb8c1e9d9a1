// Special functions for the Dirichlet-multinomial density: digamma / lgamma pieces in fp32 (the
// production arithmetic: MUFU lg2/ex2/rcp + short FMA polynomials, branch-free) and fp64 (parity build).
//
// These replace what TensorFlow's lgamma/digamma kernels and autodiff evaluate for
// tfd.DirichletMultinomial in /root/reference/sccoda/model/scCODA_model.py:743-751.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

namespace sccoda {

// ----------------------------------------------------------------------------- fp32 primitives
__device__ __forceinline__ float rcp_fast(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float lg2_fast(float x) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float ex2_fast(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
#define SCCODA_LN2F 0.6931471805599453f
#define SCCODA_LOG2EF 1.4426950408889634f

__device__ __forceinline__ float log_fast(float x) { return lg2_fast(x) * SCCODA_LN2F; }

// exp with a two-term split of log2(e) so the relative error stays ~2 ulp for |x| up to ~88
__device__ __forceinline__ float exp_acc(float x) {
    const float t = x * SCCODA_LOG2EF;
    const float n = rintf(t);
    // r = x*log2e - n computed with the product error recovered by fma
    float r = fmaf(x, SCCODA_LOG2EF, -n);
    r = fmaf(x, 1.925963033500011e-8f, r);  // low part of log2(e)
    float p = ex2_fast(r);
    // scale by 2^n via exponent arithmetic (ex2 of an integer is exact in MUFU, handles under/overflow)
    return p * ex2_fast(n);
}

// psi(x), x > 0.  x < 6: psi(x) = psi(x+6) - sum_{i<6} 1/(x+i), the sum evaluated as ONE rational
// P'(x)/P(x), P(x) = x(x+1)...(x+5) = a(a+4)(a+6), a = x(x+5).  Then the asymptotic series at z >= 6:
// psi(z) = ln z - 1/(2z) - 1/(12 z^2) + 1/(120 z^4) - 1/(252 z^6)   (next term 1/(240 z^8) < 2.5e-9).
// Branch-free (selects), 3 MUFU (lg2, rcp, rcp).
__device__ __forceinline__ float digamma_pos(float x) {
    const bool small = x < 6.0f;
    const float a = x * (x + 5.0f);
    const float P = a * (a + 4.0f) * (a + 6.0f);
    const float Nn = fmaf(fmaf(3.0f, a, 20.0f), a, 24.0f) * fmaf(2.0f, x, 5.0f);
    const float corr = small ? Nn * rcp_fast(P) : 0.0f;
    const float z = small ? x + 6.0f : x;
    const float w = rcp_fast(z);
    const float w2 = w * w;
    const float ser = w2 * fmaf(w2, fmaf(w2, -1.0f / 252.0f, 1.0f / 120.0f), -1.0f / 12.0f);
    return fmaf(lg2_fast(z), SCCODA_LN2F, fmaf(-0.5f, w, ser)) - corr;
}

// Variant for arguments known to be >= 6 (e.g. c + y with y >= 6): 2 MUFU.
__device__ __forceinline__ float digamma_big(float z) {
    const float w = rcp_fast(z);
    const float w2 = w * w;
    const float ser = w2 * fmaf(w2, fmaf(w2, -1.0f / 252.0f, 1.0f / 120.0f), -1.0f / 12.0f);
    return fmaf(lg2_fast(z), SCCODA_LN2F, fmaf(-0.5f, w, ser));
}

// Stirling remainder r(z) = lgamma(z) - [(z-1/2) ln z - z + ln(2 pi)/2], z >= 6 (error < 2e-9)
__device__ __forceinline__ float stirling_rem(float w /* = 1/z */) {
    const float w2 = w * w;
    return w * fmaf(w2, fmaf(w2, 1.0f / 1260.0f, -1.0f / 360.0f), 1.0f / 12.0f);
}

#define SCCODA_HALF_LN_2PI_F 0.9189385332046727f

// lgamma(x), x > 0, fp32: shift by 6 below 6 (minus ln P(x)), Stirling above.
__device__ __forceinline__ float lgamma_pos(float x) {
    const bool small = x < 6.0f;
    const float a = x * (x + 5.0f);
    const float P = a * (a + 4.0f) * (a + 6.0f);
    const float z = small ? x + 6.0f : x;
    const float lz = log_fast(z);
    const float w = rcp_fast(z);
    float r = fmaf(z - 0.5f, lz, -z) + SCCODA_HALF_LN_2PI_F + stirling_rem(w);
    if (small) r -= log_fast(P);
    return r;
}

// log1p(t), t >= 0, fp32: ln(1+t) by MUFU for t >= 1/2, 2 atanh(t/(2+t)) as an odd series below (rel. err ~3e-7)
__device__ __forceinline__ float log1p_pos(float t) {
    const float s = t * rcp_fast(2.0f + t);   // <= 0.2 on the series branch
    const float s2 = s * s;
    const float ser = s * fmaf(s2, fmaf(s2, fmaf(s2, fmaf(s2, 2.0f / 9.0f, 2.0f / 7.0f), 2.0f / 5.0f), 2.0f / 3.0f), 2.0f);
    return t >= 0.5f ? log_fast(1.0f + t) : ser;
}

// D(c, y) = lgamma(c + y) - lgamma(y) for a data constant y > 0 (fp32, cancellation-free):
//   y >= 6:  (y - 1/2) log1p(c/y) + c (ln(c+y) - 1) + r(c+y) - r(y)        [ycst = r(y), host fp64]
//   y <  6:  lgamma(c+y) - lgamma(y)                                          [ycst = lgamma(y), host fp64]
// The host packs ycst accordingly (sccoda_b200/_pack.py); magnitude is O(c ln(c+y)) instead of O(y ln y).
__device__ __forceinline__ float lgamma_shift_big(float c, float y, float ycst) {   // y >= 6
    const float z = c + y;
    const float w = rcp_fast(z);
    const float l1p = log1p_pos(c * rcp_fast(y));
    return fmaf(y - 0.5f, l1p, c * (log_fast(z) - 1.0f)) + (stirling_rem(w) - ycst);
}
// the same with the Stirling remainder of the count evaluated on the fly (counts >= 6 read from the items, which
// carry no host constant): r(y) costs three FMAs on top of the 1/y the form needs anyway (error < 2e-9)
__device__ __forceinline__ float lgamma_shift_big_auto(float c, float y, float* r_y = nullptr) {
    const float z = c + y;
    const float w = rcp_fast(z), wy = rcp_fast(y);
    const float l1p = log1p_pos(c * wy);
    const float ry = stirling_rem(wy);
    if (r_y) *r_y = ry;
    return fmaf(y - 0.5f, l1p, c * (log_fast(z) - 1.0f)) + (stirling_rem(w) - ry);
}
__device__ __forceinline__ float lgamma_shift_small(float c, float y, float ycst) {  // y < 6
    return lgamma_pos(c + y) - ycst;
}
__device__ __forceinline__ float lgamma_shift(float c, float y, float ycst) {
    return y >= 6.0f ? lgamma_shift_big(c, y, ycst) : lgamma_shift_small(c, y, ycst);
}

// Large concentrations (c >= SCCODA_C_JOINT): h(c, y) = lgamma(c + y) - lgamma(c) - lgamma(y) evaluated jointly,
// so that the O(c ln c) parts of lgamma(c+y) and lgamma(c) never appear (fp32 would lose them):
//   y >= 6: (c-1/2) log1p(y/c) + (y-1/2) log1p(c/y) + ln(z)/2 - ln(2 pi)/2 + r(z) - r(c) - r(y)   [ycst = r(y)]
//   y <  6: (c-1/2) log1p(y/c) + y (ln z - 1) + r(z) - r(c) - lgamma(y)                           [ycst = lgamma(y)]
#define SCCODA_C_JOINT 256.0f
__device__ __forceinline__ float lgamma_joint(float c, float y, float ycst) {
    const float z = c + y;
    const float rc = rcp_fast(c);
    const float lz = log_fast(z);
    const float a = (c - 0.5f) * log1p_pos(y * rc);
    const float rem = stirling_rem(rcp_fast(z)) - stirling_rem(rc) - ycst;
    if (y >= 6.0f) return a + fmaf(y - 0.5f, log1p_pos(c * rcp_fast(y)), fmaf(0.5f, lz, -SCCODA_HALF_LN_2PI_F)) + rem;
    return a + fmaf(y, lz - 1.0f, rem);
}

// Stirling remainder r(y) of a count >= 6 (what _pack.py stores as eycst), on the fly
__device__ __forceinline__ float stirling_rem_of(float y) { return stirling_rem(rcp_fast(y)); }

__device__ __forceinline__ float t_log1p(float v) { return log1pf(v); }
__device__ __forceinline__ double t_log1p(double v) { return log1p(v); }

__device__ __forceinline__ float sigmoid_stable(float s) {
    // == exp(s)/(1+exp(s)) (scCODA_model.py:724-725) wherever that is finite; no overflow for large |s|
    const float e = ex2_fast(-fabsf(s) * SCCODA_LOG2EF);
    const float r = rcp_fast(1.0f + e);
    return s >= 0.0f ? r : e * r;
}

// ----------------------------------------------------------------------------- fp64 (parity build)
__device__ __forceinline__ double digamma_pos(double x) {
    if (!(x > 0.0)) return (x == 0.0) ? -CUDART_INF : CUDART_NAN;
    if (isinf(x)) return x;
    double acc = 0.0;
    while (x < 10.0) {
        acc -= 1.0 / x;
        x += 1.0;
    }
    const double w = 1.0 / x, w2 = w * w;
    const double ser =
        w2 * (1.0 / 12.0 -
              w2 * (1.0 / 120.0 -
                    w2 * (1.0 / 252.0 - w2 * (1.0 / 240.0 - w2 * (1.0 / 132.0 - w2 * (691.0 / 32760.0 - w2 * (1.0 / 12.0)))))));
    return acc + log(x) - 0.5 * w - ser;
}
__device__ __forceinline__ double digamma_big(double z) { return digamma_pos(z); }

// lgamma(x), x > 0, fp64: recurrence up to x >= 10, then Stirling with 7 correction terms (|err| ~ 1e-15 rel).
// Kept out of line: it is called a handful of times per value evaluation and libdevice's lgamma is ~3 KB of code.
static __device__ __noinline__ double lgamma_f64(double x) {
    if (!(x > 0.0)) return (x == 0.0) ? CUDART_INF : CUDART_NAN;
    if (isinf(x)) return x;
    double prod = 1.0;
    while (x < 10.0) {
        prod *= x;
        x += 1.0;
    }
    const double w = 1.0 / x, w2 = w * w;
    const double ser =
        w * (1.0 / 12.0 -
             w2 * (1.0 / 360.0 -
                   w2 * (1.0 / 1260.0 -
                         w2 * (1.0 / 1680.0 - w2 * (1.0 / 1188.0 - w2 * (691.0 / 360360.0 - w2 * (1.0 / 156.0)))))));
    return (x - 0.5) * log(x) - x + 0.91893853320467274178 + ser - log(prod);
}
__device__ __forceinline__ double lgamma_pos(double x) { return lgamma_f64(x); }
// fp64: ycst = 0 and the host constant carries nothing extra — plain lgamma(c+y).
__device__ __forceinline__ double lgamma_shift(double c, double y, double ycst) { return lgamma_f64(c + y) - ycst; }
#define SCCODA_C_JOINT_F64 1e300
__device__ __forceinline__ double lgamma_joint(double c, double y, double ycst) { return lgamma_f64(c + y) - lgamma_f64(c) - ycst; }
__device__ __forceinline__ double lgamma_shift_big(double c, double y, double ycst) { return lgamma_f64(c + y) - ycst; }
__device__ __forceinline__ double stirling_rem_of(double) { return 0.0; }
__device__ __forceinline__ double lgamma_shift_big_auto(double c, double y, double* r_y = nullptr) {
    if (r_y) *r_y = 0.0;
    return lgamma_f64(c + y);
}
__device__ __forceinline__ double lgamma_shift_small(double c, double y, double ycst) { return lgamma_f64(c + y) - ycst; }

// lgamma(S) - lgamma(S + n) in fp64 for the row terms of the Dirichlet-multinomial (S = sum of concentrations,
// n = row total).  Both arguments >= 10 (the usual case): one Stirling difference with two logs; otherwise the
// general function.
static __device__ __noinline__ double lgamma_rowdiff_f64(double S, double n) {
    const double z = S + n;
    if (S >= 10.0 && z < 1e300) {
        const double ws = 1.0 / S, wz = 1.0 / z;
        const double ws2 = ws * ws, wz2 = wz * wz;
        const double rs = ws * (1.0 / 12.0 - ws2 * (1.0 / 360.0 - ws2 * (1.0 / 1260.0 - ws2 * (1.0 / 1680.0 - ws2 * (1.0 / 1188.0)))));
        const double rz = wz * (1.0 / 12.0 - wz2 * (1.0 / 360.0 - wz2 * (1.0 / 1260.0 - wz2 * (1.0 / 1680.0 - wz2 * (1.0 / 1188.0)))));
        return (S - 0.5) * log(S) - (z - 0.5) * log(z) + n + (rs - rz);
    }
    return lgamma_f64(S) - lgamma_f64(z);
}
__device__ __forceinline__ double exp_acc(double x) { return exp(x); }
__device__ __forceinline__ double sigmoid_stable(double s) {
    const double e = exp(-fabs(s));
    return s >= 0.0 ? 1.0 / (1.0 + e) : e / (1.0 + e);
}
__device__ __forceinline__ double log_fast(double x) { return log(x); }

}  // namespace sccoda
