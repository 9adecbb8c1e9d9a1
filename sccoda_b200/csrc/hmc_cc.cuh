// K2cc — HMC for case/control designs, the shape scCODA is mostly run on and the one the benchmark sweep uses:
// one covariate (D == 1), two covariate groups (Umax == 2), K <= 31 cell types, fp32, one warp per chain.
//
// Same algorithm, same Philox streams and same arithmetic per parameter as hmc_kernel (kernels_impl.cuh, which
// follows sccoda/model/scCODA_model.py:276-315, :382-424 and the TFP semantics of SURVEY §8c); what changes is
// where things live.  Lane k owns cell type k for the whole run — lane K owns the row totals, which enter the
// Dirichlet-multinomial exactly like one more cell type with concentration S_u = sum_k c_uk — and everything a
// lane owns comes in twos (the two covariate groups), so it is evaluated on packed fp32 pairs (FFMA2):
//   * alpha_k, b_offset_k, ind_raw_k, their momenta and gradients are REGISTERS (sigma_d is replicated in every
//     lane): the leapfrog update is a handful of FFMAs without shared-memory traffic;
//   * the spike-and-slab effect, both concentrations, both pair records (grouped.cuh), the items of both pairs and
//     both brackets never leave the lane: the gradient needs no shared-memory round trip and no barrier, only
//     warp shuffles (S_u, the brackets of the totals, sigma_d);
//   * shared memory holds the staged dataset at compile-time offsets (lane-major, conflict-free) and, for the
//     value pass, the concentrations.
#pragma once
#include "kernels_impl.cuh"

namespace sccoda {

// ------------------------------------------------------------------------------------------------
// Packed fp32 arithmetic: Blackwell issues two FMAs per instruction on 64-bit register pairs (FFMA2 / fma.rn.f32x2).
// ------------------------------------------------------------------------------------------------
typedef unsigned long long f32x2;

__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ f32x2 dup2(float v) { return pk2(v, v); }
__device__ __forceinline__ float lo2(f32x2 v) {
    float lo;
    float hi __attribute__((unused));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
    return lo;
}
__device__ __forceinline__ float hi2(f32x2 v) {
    float lo __attribute__((unused));
    float hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
    return hi;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 rcp2(f32x2 v) { return pk2(rcp_fast(lo2(v)), rcp_fast(hi2(v))); }
__device__ __forceinline__ f32x2 lg2_2(f32x2 v) { return pk2(lg2_fast(lo2(v)), lg2_fast(hi2(v))); }
__device__ __forceinline__ f32x2 ex2_2(f32x2 v) { return pk2(ex2_fast(lo2(v)), ex2_fast(hi2(v))); }

// Midpoint expansion of the digamma function: psi(x + 1/2) = ln x + 1/(24 x^2) - 7/(960 x^4) + O(x^-6)
// (next term 31/(8064 x^6) < 1.4e-7 for x >= 5.5).  It has no odd powers, so a count costs three packed
// operations: acc += psi(x + 1/2) - ln x with w = 1/x.
__device__ __forceinline__ f32x2 psi_mid_acc2(f32x2 w, f32x2 acc) {
    const f32x2 u = mul2(w, w);
    return fma2(u, fma2(u, dup2(-7.0f / 960.0f), dup2(1.0f / 24.0f)), acc);
}

// psi(x) on pairs for any x > 0: below 6 the recurrence psi(x) = psi(x+6) - sum_{i<6} 1/(x+i) with the sum as ONE
// rational (digamma_pos in special.cuh), then the midpoint expansion at x (+6) - 1/2
__device__ __forceinline__ f32x2 psi_any2(f32x2 x) {
    const f32x2 a = mul2(x, add2(x, dup2(5.0f)));
    const f32x2 P = mul2(mul2(a, add2(a, dup2(4.0f))), add2(a, dup2(6.0f)));
    const f32x2 Nn = mul2(fma2(fma2(dup2(3.0f), a, dup2(20.0f)), a, dup2(24.0f)), fma2(dup2(2.0f), x, dup2(5.0f)));
    const f32x2 corr = mul2(Nn, rcp2(P));
    const bool sl = lo2(x) < 6.0f, sh = hi2(x) < 6.0f;
    const f32x2 xm = add2(x, pk2(sl ? 5.5f : -0.5f, sh ? 5.5f : -0.5f));
    const f32x2 v = fma2(lg2_2(xm), dup2(SCCODA_LN2F), psi_mid_acc2(rcp2(xm), dup2(0.0f)));
    return add2(v, pk2(sl ? -lo2(corr) : 0.0f, sh ? -hi2(corr) : 0.0f));
}

// lgamma(c + y) - lgamma(y) for two counts y >= 6 at once (lgamma_shift_big of special.cuh on pairs):
//   (y - 1/2) log1p(c/y) + c (ln(c+y) - 1) + r(c+y) - r(y),   ycst = r(y) and yinv = 1/y from staging
// log1p(t), t >= 0, on pairs: ln(1+t) by MUFU for t >= 1/2, 2 atanh(t/(2+t)) as an odd series below (log1p_pos)
__device__ __forceinline__ f32x2 log1p_pos2(f32x2 t) {
    const f32x2 sq = mul2(t, rcp2(add2(t, dup2(2.0f))));
    const f32x2 s2 = mul2(sq, sq);
    f32x2 ser = fma2(s2, dup2(2.0f / 9.0f), dup2(2.0f / 7.0f));
    ser = fma2(s2, ser, dup2(2.0f / 5.0f));
    ser = fma2(s2, ser, dup2(2.0f / 3.0f));
    ser = mul2(sq, fma2(s2, ser, dup2(2.0f)));
    const f32x2 big = mul2(lg2_2(add2(t, dup2(1.0f))), dup2(SCCODA_LN2F));
    return pk2(lo2(t) >= 0.5f ? lo2(big) : lo2(ser), hi2(t) >= 0.5f ? hi2(big) : hi2(ser));
}
__device__ __forceinline__ f32x2 stirling_rem2(f32x2 w) {
    const f32x2 w2 = mul2(w, w);
    return mul2(w, fma2(w2, fma2(w2, dup2(1.0f / 1260.0f), dup2(-1.0f / 360.0f)), dup2(1.0f / 12.0f)));
}
// Two counts y >= 6 whose pairs may or may not have a concentration >= c_joint: per element either
// lgamma_shift_big (as above) or the joint form lgamma_joint of special.cuh,
//   (c-1/2) log1p(y/c) + (y-1/2) log1p(c/y) + ln(z)/2 - ln(2 pi)/2 + r(z) - r(c) - r(y),
// sharing z, 1/z, ln z, log1p(c/y) and r(z) between the two forms (the value pass of a state with a large concentration)
__device__ __forceinline__ f32x2 lgamma_mixed_big2(f32x2 c, f32x2 y, f32x2 ycst, f32x2 yinv, float c_joint) {
    const f32x2 z = add2(c, y);
    const f32x2 lz = mul2(lg2_2(z), dup2(SCCODA_LN2F));
    const f32x2 l1p = log1p_pos2(mul2(c, yinv));
    const f32x2 common = fma2(add2(y, dup2(-0.5f)), l1p, add2(stirling_rem2(rcp2(z)), mul2(ycst, dup2(-1.0f))));
    const f32x2 shift = fma2(c, add2(lz, dup2(-1.0f)), common);
    const f32x2 rc = rcp2(c);
    const f32x2 a = mul2(add2(c, dup2(-0.5f)), log1p_pos2(mul2(y, rc)));
    const f32x2 joint = add2(add2(a, fma2(lz, dup2(0.5f), dup2(-SCCODA_HALF_LN_2PI_F))), add2(common, mul2(stirling_rem2(rc), dup2(-1.0f))));
    return pk2(lo2(c) >= c_joint ? lo2(joint) : lo2(shift), hi2(c) >= c_joint ? hi2(joint) : hi2(shift));
}

__device__ __forceinline__ f32x2 lgamma_shift_big2(f32x2 c, f32x2 y, f32x2 ycst, f32x2 yinv) {
    const f32x2 z = add2(c, y);
    const f32x2 w = rcp2(z);
    const f32x2 l1p = log1p_pos2(mul2(c, yinv));
    const f32x2 cterm = fma2(mul2(c, dup2(SCCODA_LN2F)), lg2_2(z), mul2(c, dup2(-1.0f)));     // c (ln z - 1)
    const f32x2 w2 = mul2(w, w);
    const f32x2 rem = mul2(w, fma2(w2, fma2(w2, dup2(1.0f / 1260.0f), dup2(-1.0f / 360.0f)), dup2(1.0f / 12.0f)));
    return add2(fma2(add2(y, dup2(-0.5f)), l1p, cterm), add2(rem, mul2(ycst, dup2(-1.0f))));
}

// exp on pairs: one MUFU per value; 2^n is applied through the exponent field, so the argument is clamped to the
// range in which the result stays a normal number (concentrations beyond e^-86 .. e^88 are rejected states anyway:
// the value pass returns NaN above 1e6)
__device__ __forceinline__ f32x2 exp2_acc(f32x2 x) {
    const float xl = fminf(fmaxf(lo2(x), -86.0f), 88.0f), xh = fminf(fmaxf(hi2(x), -86.0f), 88.0f);
    x = pk2(xl, xh);
    const f32x2 tm = fma2(x, dup2(SCCODA_LOG2EF), dup2(12582912.0f));       // + 1.5 * 2^23: round to nearest integer
    const f32x2 n = add2(tm, dup2(-12582912.0f));
    f32x2 r = fma2(x, dup2(SCCODA_LOG2EF), mul2(n, dup2(-1.0f)));
    r = fma2(x, dup2(1.925963033500011e-8f), r);                             // low part of log2(e)
    const f32x2 p = ex2_2(r);
    const int nl = __float_as_int(lo2(tm)) - 0x4B400000, nh = __float_as_int(hi2(tm)) - 0x4B400000;
    return pk2(__int_as_float(__float_as_int(lo2(p)) + (nl << 23)), __int_as_float(__float_as_int(hi2(p)) + (nh << 23)));
}

// Two sums over the warp at once: lanes 0..15 return sum(a), lanes 16..31 return sum(b) (6 shuffles instead of 10).
__device__ __forceinline__ float warp_sum2(float a, float b, int lane) {
    const bool lo = lane < 16;
    const float recv = __shfl_xor_sync(0xffffffffu, lo ? b : a, 16);
    float v = (lo ? a : b) + recv;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_sum1(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// Two dot products of the U-turn criterion at once, in fp32: only their signs are used, and the six shuffles of
// warp_sum2 (+ one exchange) replace twenty 32-bit shuffles and ten fp64 additions of two double-precision sums.
__device__ __forceinline__ void warp_dots2(float a, float b, int lane, float& sa, float& sb) {
    const float v = warp_sum2(a, b, lane);                        // lanes 0..15: sum(a), lanes 16..31: sum(b)
    const float o = __shfl_xor_sync(0xffffffffu, v, 16);
    sa = lane < 16 ? v : o;
    sb = lane < 16 ? o : v;
}

// ------------------------------------------------------------------------------------------------
// Staging: the packed dataset (include/sccoda_b200.h) re-laid lane-major into shared memory (CcPlan, kernels.h)
// ------------------------------------------------------------------------------------------------
struct CcTile {
    uint32_t cf, yv, ov, om, rows, el;   // shared-memory byte offsets (CcPlan)
    int NT, NOT;                     // items / odd counts of the fullest pair of THIS dataset (loop bounds)
    int N, NK, nbig;
    float cnt0, cnt1;                // rows of the two covariate groups
    float c_max;                     // rejection limit of the concentrations (Dims::conc_max)
    double lconst;
};

// (group u, column) of pair index p = u*K + k (cell types) or 2K + u (row totals -> column K)
__device__ __forceinline__ void cc_pair_to_col(int p, int K, int& u, int& col) {
    if (p >= 2 * K) {
        u = p - 2 * K;
        col = K;
    } else {
        u = p >= K ? 1 : 0;
        col = p - u * K;
    }
}

__device__ __forceinline__ void cc_stage(const KArgs& a, int b, const CcPlan& pl, CcTile& t) {
    using T = float;
    const int N = a.dm.N, K = a.dm.K, NK = a.dm.NK, NP = 2 * (K + 1);
    t.cf = (uint32_t)pl.cf; t.yv = (uint32_t)pl.yv; t.ov = (uint32_t)pl.ov; t.om = (uint32_t)pl.om;
    t.rows = (uint32_t)pl.rows; t.el = (uint32_t)pl.el;
    t.N = N; t.NK = NK;
    t.c_max = (float)a.dm.conc_max;
    // [0]: items, [1]: odd counts of the fullest pair of this dataset, [2]: elements with a count >= 6
    int* bounds = sp<int>((uint32_t)pl.bounds);
    if (threadIdx.x < 2) bounds[threadIdx.x] = 0;
    const int* rstart = a.rstart + (size_t)b * 3;
    t.cnt0 = (float)(rstart[1] - rstart[0]);
    t.cnt1 = (float)(rstart[2] - rstart[1]);
    float* cf = sp<float>(t.cf);
    float* yv = sp<float>(t.yv);
    float* ov = sp<float>((uint32_t)pl.ov);
    float* om = sp<float>((uint32_t)pl.om);
    int* fill = sp<int>((uint32_t)pl.fill);
    for (int i = threadIdx.x; i < 64; i += blockDim.x) fill[i] = 0;
    const T* pcoef = static_cast<const T*>(a.pcoef) + (size_t)b * NP * 6;
    // rational coefficients: cf[j][lane] = (pair (0,lane), pair (1,lane)); lane K = row totals; lanes > K: zero
    for (int i = threadIdx.x; i < 6 * 32 * 2; i += blockDim.x) {
        const int u = i & 1, lane = (i >> 1) & 31, j = i >> 6;
        float v = 0.0f;
        if (lane <= K) v = pcoef[(lane < K ? u * K + lane : 2 * K + u) * 6 + j];
        cf[i] = v;
    }
    // item counts: padding is the count 6, which contributes psi(c+6) - psi(c+6) = 0; odd counts: padding has the
    // multiplicity 0
    for (int i = threadIdx.x; i < pl.NTC * kItemR * 64; i += blockDim.x) yv[i] = 6.0f;
    for (int i = threadIdx.x; i < a.NOT * 64; i += blockDim.x) {
        ov[i] = 1.0f;
        om[i] = 0.0f;
    }
    __syncthreads();
    const T* iy = static_cast<const T*>(a.iy) + (size_t)b * kItemR * a.NImax;
    const uint16_t* ipair = a.ipair + (size_t)b * a.NImax;
    const uint16_t* ipos = a.ipos + (size_t)b * a.NImax;
    const int NI = a.NI[b];
    for (int i = threadIdx.x; i < NI * kItemR; i += blockDim.x) {
        const int j = i / NI, it = i - j * NI;
        const int p = ipair[it], tt = ipos[it] - p * a.NF;
        int u, col;
        cc_pair_to_col(p, K, u, col);
        if (tt < 0 || tt >= a.NT || col > K || u > 1) continue;   // a layout that contradicts its own header: never write outside the tile
        yv[((tt * kItemR + j) * 32 + col) * 2 + u] = iy[j * a.NImax + it];
        if (j == 0) {
            atomicMax(&bounds[0], tt + 1);
            atomicMax(&fill[col * 2 + u], tt * kItemR + (int)a.inreal[(size_t)b * a.NImax + it]);   // data counts of the pair
        }
    }
    __syncthreads();
    // Odd counts y (below 6, not in {0..5}: the 0.5 pseudocount, non-integer data):
    //   psi(c + y) - psi(c + 6) = [psi(c + y + 6) - psi(c + 6)] - sum_{i<6} 1/(c + y + i)
    // The bracket has the form of an item count, so y + 6 goes into the next free item slot of the pair (usually
    // padding of its last item: no extra work); the sum is a rational in c + y, evaluated once per DISTINCT odd value
    // of the pair and weighted with how often it occurs (ov / om; cc_grad).
    const int NG = a.NG[b];
    const T* oy = static_cast<const T*>(a.oy) + (size_t)b * a.NOmax;
    for (int q = threadIdx.x; q < NG; q += blockDim.x) {
        const uint32_t od = a.odesc[(size_t)b * a.NGmax + q];
        int u, col;
        cc_pair_to_col(a.opair[(size_t)b * a.NGmax + q], K, u, col);
        if (col > K || u > 1) continue;
        const int base = fill[col * 2 + u];
        int n = (int)(od >> 16);
        if (n > a.NOT) n = a.NOT;                                // (the header's bounds are what the tile was sized for)
        int nd = 0;
        for (int tt = 0; tt < n; ++tt) {
            const float y = oy[(od & 0xffffu) + tt];
            const int slot = base + tt;
            if (slot >= pl.NTC * kItemR) break;
            yv[(slot * 32 + col) * 2 + u] = y + 6.0f;            // slot = item * R + position
            int d = 0;
            while (d < nd && ov[(d * 32 + col) * 2 + u] != y) ++d;
            if (d == nd) {
                ov[(d * 32 + col) * 2 + u] = y;
                ++nd;
            }
            om[(d * 32 + col) * 2 + u] += 1.0f;
        }
        if (n > 0) atomicMax(&bounds[0], (base + n - 1) / kItemR + 1);
        atomicMax(&bounds[1], nd);
    }
    // value pass, rows: (total n, r(n) or lgamma(n) as in _pack.py's eycst, group); the row terms are evaluated as
    // -[lgamma(S+n) - lgamma(S) - lgamma(n)] in fp32 (cancellation-free), so sum_n lgamma(n) moves into the constant
    float4* rows = sp<float4>(t.rows);
    const T* ntot = static_cast<const T*>(a.ntot) + (size_t)b * N;
    const int* grp = a.grp + (size_t)b * N;
    double lg_sum = 0.0;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        const double nn = (double)ntot[n];
        const double lg = lgamma(nn);
        const double rc = nn >= 6.0 ? lg - ((nn - 0.5) * log(nn) - nn + 0.91893853320467274178) : lg;
        rows[n] = make_float4((float)nn, (float)rc, __int_as_float(grp[n]), 0.0f);
        lg_sum += lg;
    }
    double* red = sp<double>((uint32_t)pl.red);
    __syncthreads();
    for (int o = 16; o > 0; o >>= 1) lg_sum += __shfl_xor_sync(0xffffffffu, lg_sum, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = lg_sum;   // one slot per warp (up to 32): deterministic sum below
    // value pass, elements: (count, host constant, concentration slot u*32 + k), counts >= 6 first.  The partition
    // is done by warp 0 with ballots, in list order, so that the summation order never depends on the launch.
    float4* el = sp<float4>(t.el);
    const T* ey = static_cast<const T*>(a.ey) + (size_t)b * NK;
    const T* eycst = static_cast<const T*>(a.eycst) + (size_t)b * NK;
    const uint32_t* eidx = a.eidx + (size_t)b * NK;
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        int nbig = 0;
        for (int base = 0; base < NK; base += 32) nbig += __popc(__ballot_sync(0xffffffffu, base + lane < NK && ey[base + lane] >= 6.0f));
        int ib = 0, is = nbig;
        for (int base = 0; base < NK; base += 32) {
            const int e = base + lane;
            const bool in = e < NK;
            const float y = in ? ey[e] : 0.0f;
            const bool big = in && y >= 6.0f;
            const uint32_t mb = __ballot_sync(0xffffffffu, big), ms = __ballot_sync(0xffffffffu, in && !big);
            const uint32_t lt = (1u << lane) - 1u;
            if (in) {
                const int uk = (int)(eidx[e] & 0xffffu);
                const int u = uk >= K ? 1 : 0;
                const int dst = big ? ib + __popc(mb & lt) : is + __popc(ms & lt);
                el[dst] = make_float4(y, eycst[e], __int_as_float(u * 32 + (uk - u * K)), 1.0f / y);
            }
            ib += __popc(mb);
            is += __popc(ms);
        }
        if (lane == 0) bounds[2] = nbig;
    }
    __syncthreads();
    t.NT = bounds[0];
    t.NOT = bounds[1];
    t.nbig = bounds[2];
    double lg_all = 0.0;
    for (int w = 0; w < (int)((blockDim.x + 31) >> 5); ++w) lg_all += red[w];   // rows are dealt to the threads in order: only
    t.lconst = a.lconst[b] - lg_all;                                            // the first ceil(N / 32) warps hold a part
}

// ------------------------------------------------------------------------------------------------
// Gradient
// ------------------------------------------------------------------------------------------------
// The four parameters a lane owns: alpha_k, b_offset_k, ind_raw_k (zero on the reference lane and on lanes >= K)
// and sigma_d (uniform).
struct CcVec {
    float a, b, i, s;
};

struct CcLane {
    int K;
    bool has, active;    // lane < K; lane < K and lane != ref
    float x0, x1;        // the covariate value of the two groups
};

// sum_i W_i/(c+i) with ONE reciprocal: the three quadratics c(c+5), (c+1)(c+4), (c+2)(c+3) over their product
// (overflows for c > ~2e6, where states are rejected anyway); cf = the lane's six packed coefficients
__device__ __forceinline__ f32x2 pair_rational2(f32x2 c, const f32x2* __restrict__ cf) {
    const f32x2 a = mul2(c, add2(c, dup2(5.0f)));
    const f32x2 a4 = add2(a, dup2(4.0f)), a6 = add2(a, dup2(6.0f));
    const f32x2 a46 = mul2(a4, a6);
    const f32x2 n0 = fma2(cf[0 * 32], c, cf[1 * 32]), n1 = fma2(cf[2 * 32], c, cf[3 * 32]), n2 = fma2(cf[4 * 32], c, cf[5 * 32]);
    const f32x2 t = fma2(n1, a6, mul2(n2, a4));
    const f32x2 num = fma2(a, t, mul2(n0, a46));
    return mul2(num, rcp2(mul2(a, a46)));
}

// -sum_{i<6} 1/(z + i) = -(2z + 5) [1/a + 1/(a+4) + 1/(a+6)], a = z (z + 5), with one reciprocal
__device__ __forceinline__ f32x2 r6_neg2(f32x2 z) {
    const f32x2 a = mul2(z, add2(z, dup2(5.0f)));
    const f32x2 a4 = add2(a, dup2(4.0f)), a6 = add2(a, dup2(6.0f));
    const f32x2 a46 = mul2(a4, a6);
    const f32x2 num = fma2(a, add2(a4, a6), a46);
    return mul2(mul2(fma2(dup2(-2.0f), z, dup2(-5.0f)), num), rcp2(mul2(a, a46)));
}

// Two items at once: sum_j [psi(c + y_j) - psi(c + 6)] for the R = 5 counts of each (all >= 6), with the midpoint
// expansion at x_j = c + y_j - 1/2 (cm = c - 1/2); a = w^5 and nb = -5 (psi(x+1/2) - ln x) at x = c + 5.5, w = 1/x
// refer the log-product and the series to c + 6 (cf. PairRec / item_eval in grouped.cuh).
// Reciprocals: one MUFU for the first four counts (1/x_j = product of the others / product of all), one for the
// fifth; one lg2 of the product ratio.
__device__ __forceinline__ f32x2 item_eval2(f32x2 cm, f32x2 a, f32x2 nb, const f32x2* __restrict__ y) {
    static_assert(SCCODA_ITEM_R == 5, "item_eval2 is written for 5 counts per item");
    f32x2 z[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) z[j] = add2(cm, y[j * 32]);
    const f32x2 p01 = mul2(z[0], z[1]), p23 = mul2(z[2], z[3]);
    const f32x2 r = rcp2(mul2(p01, p23));
    const f32x2 r01 = mul2(p23, r), r23 = mul2(p01, r);
    f32x2 acc = nb;
    acc = psi_mid_acc2(mul2(z[1], r01), acc);
    acc = psi_mid_acc2(mul2(z[0], r01), acc);
    acc = psi_mid_acc2(mul2(z[3], r23), acc);
    acc = psi_mid_acc2(mul2(z[2], r23), acc);
    acc = psi_mid_acc2(rcp2(z[4]), acc);
    const f32x2 prod = mul2(mul2(p01, a), mul2(p23, z[4]));
    return fma2(lg2_2(prod), dup2(SCCODA_LN2F), acc);
}

struct CcGrad {
    CcVec g;
    f32x2 c;   // concentrations of the lane's two pairs (lane K: S_0, S_1)
};

// Gradient of the log-density at q in the pair / item formulation of grouped.cuh.
__device__ __forceinline__ CcGrad cc_grad(const CcTile& t, const CcLane& ln, const CcVec& q, int lane) {
    // spike-and-slab effect of the lane's cell type                                             (:724-734)
    const float id = ln.active ? sigmoid_stable(50.0f * q.i) : 0.0f;
    const float bet = id * q.s * q.b;
    // concentrations of the two groups (:743); lane K takes the row sums S_u = sum_k c_uk
    f32x2 c = exp2_acc(fma2(pk2(ln.x0, ln.x1), dup2(bet), dup2(q.a)));
    const float c0 = ln.has ? lo2(c) : 0.0f, c1 = ln.has ? hi2(c) : 0.0f;
    const float S = warp_sum2(c0, c1, lane);                                  // lanes 0..15: S_0, lanes 16..31: S_1
    const float So = __shfl_xor_sync(0xffffffffu, S, 16);
    if (lane == ln.K) c = lane < 16 ? pk2(S, So) : pk2(So, S);
    // reference point of the items: x6 = c + 5.5, w6 = 1/x6, a = w6^5, nb = -5 (psi(c+6) - ln x6)
    const f32x2* cf = sp<f32x2>(t.cf) + lane;
    const f32x2 cm = add2(c, dup2(-0.5f));
    const f32x2 w6 = rcp2(add2(c, dup2(5.5f)));
    const f32x2 w62 = mul2(w6, w6);
    const f32x2 a5 = mul2(mul2(w62, w62), w6);
    const f32x2 nb = mul2(psi_mid_acc2(w6, dup2(0.0f)), dup2(-(float)SCCODA_ITEM_R));
    // brackets Br = rational sum_i W_i/(c+i) + items + odd counts, both groups at once
    f32x2 br = pair_rational2(c, cf);
    const f32x2* yv = sp<f32x2>(t.yv) + lane;
#pragma unroll 1
    for (int tt = 0; tt < t.NT; ++tt, yv += SCCODA_ITEM_R * 32) br = add2(br, item_eval2(cm, a5, nb, yv));
    // odd counts: their shifted values sit in the items above; what is left is -m sum_{i<6} 1/(c + y + i) per distinct
    // odd value y of the pair (m = its multiplicity; cc_stage)
    yv = sp<f32x2>(t.ov) + lane;
    const f32x2* mv = sp<f32x2>(t.om) + lane;
#pragma unroll 1
    for (int tt = 0; tt < t.NOT; ++tt, yv += 32, mv += 32) br = fma2(*mv, r6_neg2(add2(c, *yv)), br);
    // d logp / d log c_uk = c_uk (Br_uk - Br_total(u)), chain rule
    const float bt0 = __shfl_sync(0xffffffffu, lo2(br), ln.K), bt1 = __shfl_sync(0xffffffffu, hi2(br), ln.K);
    const float G0 = c0 * (lo2(br) - bt0), G1 = c1 * (hi2(br) - bt1);        // c0 = c1 = 0 on lanes >= K
    CcGrad out;
    out.c = c;
    out.g.a = fmaf(q.a, -0.04f, G0 + G1);                                      // d/d alpha_k
    const float tg = fmaf(ln.x0, G0, ln.x1 * G1) * id;                         // id == 0 on inactive lanes
    out.g.b = fmaf(tg, q.s, -q.b);                                             // d/d b_offset_k
    out.g.i = fmaf(tg * q.s * q.b, 50.0f * (1.0f - id), -q.i);                 // d/d ind_raw_k
    // d/d sigma_d: sum_k g_k ind_k b_offset_k - HalfCauchy term (zero prior gradient below 0 [TFP])
    out.g.s = warp_sum1(tg * q.b) - (q.s >= 0.0f ? 2.0f * q.s * rcp_fast(fmaf(q.s, q.s, 1.0f)) : 0.0f);
    return out;
}

// ------------------------------------------------------------------------------------------------
// Value of the log-density (same terms, guards and precision strategy as dm_value in model.cuh)
// ------------------------------------------------------------------------------------------------
static __device__ __noinline__ double cc_value(const CcTile& t, uint32_t pc_off, bool has, float qa, float qb, float qi, float qs,
                                        f32x2 c, int K, int lane) {
    float* pc = sp<float>(pc_off);
    const float c0 = lo2(c), c1 = hi2(c);
    __syncwarp();
    pc[lane] = c0;            // slot u*32 + column; column K holds S_u
    pc[32 + lane] = c1;
    __syncwarp();
    double lp = 0.0;
    // priors (:703-741); the theta-independent constants live in lconst.  Parameters a lane does not own are zero.
    lp += (double)fmaf(qa * qa, -0.02f, -0.5f * fmaf(qb, qb, qi * qi));
    if (lane == 0 && qs >= 0.0f) lp -= (double)log1pf(qs * qs);
    // DirichletMultinomial [TFP]: lbeta(c+y) - lbeta(c)                               (:746-751)
    // Guard (numerics): beyond c_max the lgamma differences have lost all precision => NaN => rejected.
    const float c_max = t.c_max, c_joint = SCCODA_C_JOINT;
    bool need_fix = false, bad = false;
    if (has) {
        float acc = 0.0f;
        if (c0 < c_joint) acc = -t.cnt0 * lgamma_pos(c0); else need_fix = true;
        if (c1 < c_joint) acc = fmaf(-t.cnt1, lgamma_pos(c1), acc); else need_fix = true;
        bad = !(c0 <= c_max) || !(c1 <= c_max);
        lp += (double)acc;
    }
    if (bad) lp = CUDART_NAN;
    {
        // rows: -[lgamma(S_u + n) - lgamma(S_u) - lgamma(n)], joint form for S_u >= 6 (the usual case)
        const float S0 = pc[K], S1 = pc[32 + K];
        const float4* rows = sp<float4>(t.rows);
        float acc = 0.0f;
#pragma unroll 1
        for (int n = lane; n < t.N; n += 32) {
            const float4 r = rows[n];
            const float S = __float_as_int(r.z) ? S1 : S0;
            acc -= S >= 6.0f ? lgamma_joint(S, r.x, r.y) : lgamma_shift(S, r.x, r.y) - lgamma_pos(S);
        }
        lp += (double)acc;
    }
    {
        // When some concentration is >= c_joint (warp-uniform flag), the elements of those pairs take the joint
        // cancellation-free form, which also carries the -lgamma(c); otherwise the packed fast loop.
        const bool fix = __any_sync(0xffffffffu, need_fix);
        const float4* el = sp<float4>(t.el);
        float acc = 0.0f;
        if (!fix) {
            // counts >= 6, two per lane and iteration on packed pairs; a last unpaired one alone
            f32x2 acc2 = dup2(0.0f);
            int e = lane;
#pragma unroll 1
            for (; e + 32 < t.nbig; e += 64) {
                const float4 va = el[e], vb = el[e + 32];
                acc2 = add2(acc2, lgamma_shift_big2(pk2(pc[__float_as_int(va.z)], pc[__float_as_int(vb.z)]), pk2(va.x, vb.x),
                                                    pk2(va.y, vb.y), pk2(va.w, vb.w)));
            }
            acc = lo2(acc2) + hi2(acc2);
            if (e < t.nbig) {
                const float4 v = el[e];
                acc += lgamma_shift_big(pc[__float_as_int(v.z)], v.x, v.y);
            }
#pragma unroll 1
            for (e = t.nbig + lane; e < t.NK; e += 32) {
                const float4 v = el[e];
                acc += lgamma_shift(pc[__float_as_int(v.z)], v.x, v.y);
            }
        } else {
            // counts >= 6 two at a time, the form chosen per element; the rest one by one
            f32x2 acc2 = dup2(0.0f);
            int e = lane;
#pragma unroll 1
            for (; e + 32 < t.nbig; e += 64) {
                const float4 va = el[e], vb = el[e + 32];
                acc2 = add2(acc2, lgamma_mixed_big2(pk2(pc[__float_as_int(va.z)], pc[__float_as_int(vb.z)]), pk2(va.x, vb.x),
                                                    pk2(va.y, vb.y), pk2(va.w, vb.w), c_joint));
            }
            acc = lo2(acc2) + hi2(acc2);
            if (e < t.nbig) {
                const float4 v = el[e];
                const float cc = pc[__float_as_int(v.z)];
                acc += cc >= c_joint ? lgamma_joint(cc, v.x, v.y) : lgamma_shift_big(cc, v.x, v.y);
            }
#pragma unroll 1
            for (e = t.nbig + lane; e < t.NK; e += 32) {
                const float4 v = el[e];
                const float cc = pc[__float_as_int(v.z)];
                acc += cc >= c_joint ? lgamma_joint(cc, v.x, v.y) : lgamma_shift(cc, v.x, v.y);
            }
        }
        lp += (double)acc;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lp += __shfl_xor_sync(0xffffffffu, lp, o);
    lp += t.lconst;
    if (qs < 0.0f) lp = -CUDART_INF;   // HalfCauchy support [TFP]
    return lp;
}

#ifndef SCCODA_CC_HELPERS_ONLY   // hmc_mg.cuh reuses the helpers above, not the kernels
// ------------------------------------------------------------------------------------------------
// Kernel
// ------------------------------------------------------------------------------------------------
// The body is shared by two entry points that differ in their register budget only:
//   hmc_cc_kernel      72 registers, 7 CTAs of 128 threads per SM: every chain of a full sweep (>= ~2400 chains) resident
//   hmc_cc_wide_kernel the same budget as ONE CTA of up to 28 warps per SM (see WIDE below): the sweep's launch
//   hmc_cc_lat_kernel, hmc_cc_lat16_kernel  no register cap / 128 instead of 72 registers: for launches that leave the
//                      device under-filled (at most 12 / 16 warps per SM, e.g. a shard of the sweep on each of 8 GPUs:
//                      512 chains on 592 schedulers), where a warp runs at the latency of its own dependency chains and
//                      the spills / re-materialisations of the capped build are pure loss
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned sm_id() {
    unsigned v;
    asm volatile("mov.u32 %0, %smid;" : "=r"(v));
    return v;
}

// WIDE: one CTA per SM that holds a balanced share of ALL chains of the launch (a.wide = the largest share: 28 warps
// for 4096 chains on 148 SMs; the chains are dealt out by their global index, the CTA stages every dataset its chains
// belong to) and meets at a CTA barrier every 16 transitions.  The warp schedulers are not fair: in the narrow launch
// (7 independent CTAs per SM) the chains of one SM finish between 30 and 59 ms of a 59 ms launch, so the SM runs its
// last third with a fraction of its warps and little latency hiding left; chains that advance together all finish
// together (49.8 ms for the same launch).
template <bool WIDE>
__device__ __forceinline__ void hmc_cc_body(const KArgs& a) {
    using T = float;
    const unsigned long long t_start = global_timer_ns();
    const int lane = threadIdx.x & 31, gid = threadIdx.x >> 5;
    const int K = a.dm.K, K1 = a.dm.K1, P = a.dm.P;
    CcPlan pl = plan_cc(a.dm.N, K, a.NT, a.NOT, 0);
    CcTile tile;
    int b, c;
    size_t data_bytes = pl.data_total;
    if constexpr (WIDE) {
        const long long n_chains = (long long)a.B * a.C;
        const int lo = (int)(n_chains * blockIdx.x / gridDim.x), hi = (int)(n_chains * (blockIdx.x + 1) / gridDim.x);
        const int chain_g = lo + gid;                    // global chain of this warp (>= hi: nothing to do)
        const int b_lo = lo / a.C, b_hi = (hi - 1) / a.C;
        b = chain_g / a.C;
        c = chain_g - b * a.C;
        data_bytes = (size_t)wide_datasets(a.wide, a.C) * pl.data_total;
#pragma unroll 1
        for (int bb = b_lo; bb <= b_hi; ++bb) {          // the whole CTA stages one dataset after the other
            CcPlan pd = pl;
            const size_t sh = (size_t)(bb - b_lo) * pl.data_total;
            pd.cf += sh; pd.yv += sh; pd.ov += sh; pd.om += sh; pd.rows += sh; pd.el += sh; pd.bounds += sh; pd.red += sh; pd.fill += sh;
            CcTile td;
            cc_stage(a, bb, pd, td);
            if (bb == b) tile = td;
        }
        if (chain_g >= hi) return;
    } else {
        const int cpc = blockDim.x >> 5;
        const int ctas_per_ds = (a.C + cpc - 1) / cpc;
        b = blockIdx.x / ctas_per_ds;
        c = (blockIdx.x - b * ctas_per_ds) * cpc + gid;
        cc_stage(a, b, pl, tile);
        if (c >= a.C) return;
    }
    const uint32_t chain_off = (uint32_t)(data_bytes + (size_t)gid * pl.chain_total);
    const uint32_t pc_off = chain_off + (uint32_t)pl.pc;
    T* const vec = sp<T>(chain_off + (uint32_t)pl.vec);   // one flat state vector: momentum transpose, carry in / out
    const size_t chain = (size_t)b * a.C + c;
    const RngKey key{(uint32_t)(a.seed & 0xffffffffu), (uint32_t)(a.seed >> 32), a.chain_id0 + (uint32_t)chain};
    const int ref = a.ref[b];

    CcLane ln;
    ln.K = K;
    ln.has = lane < K;
    ln.active = ln.has && lane != ref;
    {
        const T* xu = static_cast<const T*>(a.xu) + (size_t)b * 2;
        ln.x0 = xu[0];
        ln.x1 = a.U[b] > 1 ? xu[1] : T(0);
    }
    // flat indices of the lane's parameters (order of scCODA_model.py:692-693 with D == 1)
    const int jj = lane - (lane > ref);
    const int ix_b = 1 + jj, ix_i = 1 + K1 + jj, ix_a = 1 + 2 * K1 + lane;

    auto gather = [&](const T* v) {               // flat shared-memory vector -> lane registers
        CcVec r;
        r.s = v[0];
        r.a = ln.has ? v[ix_a] : T(0);
        r.b = ln.active ? v[ix_b] : T(0);
        r.i = ln.active ? v[ix_i] : T(0);
        return r;
    };
    auto scatter = [&](T* v, const CcVec& r) {    // lane registers -> flat vector (shared or global)
        if (lane == 0) v[0] = r.s;
        if (ln.has) v[ix_a] = r.a;
        if (ln.active) {
            v[ix_b] = r.b;
            v[ix_i] = r.i;
        }
    };

    Adapt<T> ad;
    load_chain_state<T>(a, chain, vec, ad, lane, 32);
    __syncwarp();
    CcVec thc = gather(vec);                       // current state
    CcVec gc = {T(0), T(0), T(0), T(0)};           // its gradient
    const int L = a.num_leapfrog;
    const double log_target = log(a.target_accept);
    T* draws = static_cast<T*>(a.draws) + chain * (size_t)a.num_results * P;
    T* stats = static_cast<T*>(a.stats) + chain * (size_t)a.num_results * SCCODA_NSTAT;

    // Transition t = step_begin - 1 is a pseudo-step that only evaluates the starting point (value + gradient)
    // and "accepts" it: the hot loop contains one gradient and one value call site.
    double lp_cur = 0.0;
#pragma unroll 1
    for (int t = a.step_begin - 1; t < a.step_end; ++t) {
        const bool boot = t < a.step_begin;
        const T eps = ad.eps;
        T kin = T(0);
        CcVec q = thc, m = {T(0), T(0), T(0), T(0)};
        if (!boot) {
            // momentum draw in the flat parameter order (the stream contract of philox.cuh), regrouped by cell type
            __syncwarp();
#pragma unroll 1
            for (int i = lane; i < P; i += 32) {
                const T p = frozen_param(a.dm, i) ? T(0) : rng_normal<T>(key, (uint32_t)i, (uint32_t)t);
                kin += p * p;
                vec[i] = p;
            }
            __syncwarp();
            const CcVec p = gather(vec);
            // first half kick and first drift: q1 = q0 + eps (p + eps/2 grad(q0))
            const T h = T(0.5) * eps;
            m.a = fmaf(h, gc.a, p.a); m.b = fmaf(h, gc.b, p.b); m.i = fmaf(h, gc.i, p.i); m.s = fmaf(h, gc.s, p.s);
            q.a = fmaf(eps, m.a, thc.a); q.b = fmaf(eps, m.b, thc.b); q.i = fmaf(eps, m.i, thc.i); q.s = fmaf(eps, m.s, thc.s);
        }
        const int nl = boot ? 1 : L;
        CcGrad gr;
#pragma unroll 1
        for (int l = 0; l < nl; ++l) {
            gr = cc_grad(tile, ln, q, lane);
            if (a.dm.freeze) gr.g.s = gr.g.i = T(0);   // SimpleModel: sigma_d and ind_raw stay where they are
            if (!boot) {
                const bool last = (l == nl - 1);
                const T f = last ? T(0.5) * eps : eps;
                m.a = fmaf(f, gr.g.a, m.a); m.b = fmaf(f, gr.g.b, m.b); m.i = fmaf(f, gr.g.i, m.i); m.s = fmaf(f, gr.g.s, m.s);
                if (!last) {
                    q.a = fmaf(eps, m.a, q.a); q.b = fmaf(eps, m.b, q.b); q.i = fmaf(eps, m.i, q.i); q.s = fmaf(eps, m.s, q.s);
                }
            }
        }
        // value at the end point, from the concentrations of the last gradient evaluation
        const double lp_new = cc_value(tile, pc_off, ln.has, q.a, q.b, q.i, q.s, gr.c, K, lane);
        bool acc = true;
        double lar = 0.0;
        if (!boot) {
            // kinetic energies: the momenta of sigma_d are replicated, count them once
            kin -= fmaf(m.a, m.a, fmaf(m.b, m.b, m.i * m.i));
            if (lane == 0) kin -= m.s * m.s;
            double dkin = (double)kin;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) dkin += __shfl_xor_sync(0xffffffffu, dkin, o);
            // [TFP] log_accept_ratio = safe_sum([logp', -logp, kinetic - kinetic']); non-finite -> -inf
            lar = lp_new - lp_cur + 0.5 * dkin;
            if (!isfinite(lar)) lar = -CUDART_INF;
            const T u = rng_uniform<T>(key, 0u, (uint32_t)t, STREAM_ACCEPT);
            acc = (double)logf((float)u) < lar;   // u is a 24-bit uniform, exact in fp32
        }
        if (acc) {
            thc = q;
            gc = gr.g;
            lp_cur = lp_new;
        }
        if (boot) continue;
        if constexpr (WIDE) {
            if ((t & 15) == 0) __syncthreads();   // the chains of the SM advance together (exited warps do not count)
        }
        T traced_eps = eps;
        if (a.method == SCCODA_METHOD_HMC) {
            ad.simple(lar, t, a.num_adapt, log_target);
        } else {
            ad.dual(fmin(lar, 0.0), a.num_adapt, (T)a.target_accept);
            traced_eps = t_exp(ad.log_avg);  // scCODA_model.py:408,419
        }
        const int r = t - a.num_burnin;
        if (r >= 0 && a.store_draws) {
            scatter(draws + (size_t)r * P, thc);
            if (lane == 0)
                write_stats<T>(stats + (size_t)r * SCCODA_NSTAT, lp_cur, lar, acc, traced_eps, lar < -1000.0, L,
                               CUDART_NAN, false);
        }
    }
    __syncwarp();
    scatter(vec, thc);
    __syncwarp();
    store_chain_state<T>(a, chain, vec, ad, lp_cur, lane, 32);
    // diagnostics in the spare carry slots: how long this chain ran (ns, global timer) and on which SM
    if (a.carry != nullptr && lane == 0) {
        double* cr = a.carry + chain * (P + SCCODA_NCARRY);
        cr[P + 5] = (double)(global_timer_ns() - t_start);
        cr[P + 6] = (double)sm_id();
    }
}

__global__ void __launch_bounds__(128, 7) hmc_cc_kernel(const __grid_constant__ KArgs a) { hmc_cc_body<false>(a); }
__global__ void __launch_bounds__(896, 1) hmc_cc_wide_kernel(const __grid_constant__ KArgs a) { hmc_cc_body<true>(a); }
// the wide geometry for 13 ... 16 chains per SM: half the warps, 128 registers
__global__ void __launch_bounds__(512, 1) hmc_cc_wide16_kernel(const __grid_constant__ KArgs a) { hmc_cc_body<true>(a); }

// K1cc — value and gradient at given states through cc_grad / cc_value, i.e. exactly what the sampler above evaluates
// at every leapfrog (sccoda_logp_grad_as; target_log_prob_fn + autodiff of scCODA_model.py:756-760): theta[B,C,P] in
// a.init, gradient to a.draws, value to a.carry[chain].
__global__ void __launch_bounds__(128) logp_grad_cc_kernel(const __grid_constant__ KArgs a) {
    using T = float;
    const int lane = threadIdx.x & 31, gid = threadIdx.x >> 5;
    const int cpc = blockDim.x >> 5;
    const int ctas_per_ds = (a.C + cpc - 1) / cpc;
    const int b = blockIdx.x / ctas_per_ds;
    const int c = (blockIdx.x - b * ctas_per_ds) * cpc + gid;
    const int K = a.dm.K, K1 = a.dm.K1, P = a.dm.P;
    const CcPlan pl = plan_cc(a.dm.N, K, a.NT, a.NOT, 0);
    CcTile tile;
    cc_stage(a, b, pl, tile);
    if (c >= a.C) return;
    const uint32_t chain_off = (uint32_t)(pl.data_total + (size_t)gid * pl.chain_total);
    const uint32_t pc_off = chain_off + (uint32_t)pl.pc;
    T* const vec = sp<T>(chain_off + (uint32_t)pl.vec);
    const size_t chain = (size_t)b * a.C + c;
    const int ref = a.ref[b];
    CcLane ln;
    ln.K = K;
    ln.has = lane < K;
    ln.active = ln.has && lane != ref;
    {
        const T* xu = static_cast<const T*>(a.xu) + (size_t)b * 2;
        ln.x0 = xu[0];
        ln.x1 = a.U[b] > 1 ? xu[1] : T(0);
    }
    const int jj = lane - (lane > ref);
    const int ix_b = 1 + jj, ix_i = 1 + K1 + jj, ix_a = 1 + 2 * K1 + lane;
    const T* theta = static_cast<const T*>(a.init) + chain * P;
    for (int i = lane; i < P; i += 32) vec[i] = theta[i];
    __syncwarp();
    CcVec q;
    q.s = vec[0];
    q.a = ln.has ? vec[ix_a] : T(0);
    q.b = ln.active ? vec[ix_b] : T(0);
    q.i = ln.active ? vec[ix_i] : T(0);
    const CcGrad gr = cc_grad(tile, ln, q, lane);
    const double lp = cc_value(tile, pc_off, ln.has, q.a, q.b, q.i, q.s, gr.c, K, lane);
    T* out = static_cast<T*>(a.draws) + chain * P;
    if (lane == 0) out[0] = gr.g.s;
    if (ln.has) out[ix_a] = gr.g.a;
    if (ln.active) {
        out[ix_b] = gr.g.b;
        out[ix_i] = gr.g.i;
    }
    if (lane == 0) a.carry[chain] = lp;
}
// (128, 4): at most 128 registers, so that 16 warps per SM are resident at once; (128, 1): no cap, up to 12 warps per SM
__global__ void __launch_bounds__(128, 4) hmc_cc_lat16_kernel(const __grid_constant__ KArgs a) { hmc_cc_body<false>(a); }
__global__ void __launch_bounds__(128, 1) hmc_cc_lat_kernel(const __grid_constant__ KArgs a) { hmc_cc_body<false>(a); }

// ------------------------------------------------------------------------------------------------
// K3cc — NUTS for the same case/control shape: the iterative tree of nuts_kernel (kernels_impl.cuh; [TFP]
// NoUTurnSampler.one_step as restated in SURVEY §8c: multinomial sampling, generalized U-turn, checkpoint memory, dual
// averaging) on the lane-per-cell-type state.  The leaf being integrated (q, p, gradient, running momentum sum) lives
// in registers; the tree's stored vectors are float4 slots per lane in shared memory (one LDS.128 / STS.128 each).
// Same Philox streams and the same arithmetic per parameter as the generic kernel.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ CcVec cc_ld(const float4* slot) {
    const float4 v = *slot;
    return CcVec{v.x, v.y, v.z, v.w};
}
__device__ __forceinline__ void cc_st(float4* slot, const CcVec& v) { *slot = make_float4(v.a, v.b, v.i, v.s); }
// per-lane part of a dot product over the flat parameter vector (sigma_d is replicated: lane 0 counts it)
__device__ __forceinline__ float cc_dot(const CcVec& x, const CcVec& y, int lane) {
    float d = fmaf(x.a, y.a, fmaf(x.b, y.b, x.i * y.i));
    if (lane == 0) d = fmaf(x.s, y.s, d);
    return d;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(128, 4) nuts_cc_kernel(const __grid_constant__ KArgs a) {
    using T = float;
    const int lane = threadIdx.x & 31, gid = threadIdx.x >> 5;
    const int cpc = blockDim.x >> 5;
    const int ctas_per_ds = (a.C + cpc - 1) / cpc;
    const int b = blockIdx.x / ctas_per_ds;
    const int c = (blockIdx.x - b * ctas_per_ds) * cpc + gid;
    const int K = a.dm.K, K1 = a.dm.K1, P = a.dm.P, maxd = a.max_depth;
    const CcPlan pl = plan_cc(a.dm.N, K, a.NT, a.NOT, 11 + 2 * (maxd + 1));
    CcTile tile;
    cc_stage(a, b, pl, tile);
    if (c >= a.C) return;
    const uint32_t chain_off = (uint32_t)(pl.data_total + (size_t)gid * pl.chain_total);
    const uint32_t pc_off = chain_off + (uint32_t)pl.pc;
    T* const vec = sp<T>(chain_off + (uint32_t)pl.vec);
    float4* const slots = sp<float4>(chain_off + (uint32_t)pl.slots) + lane;   // slot k of this lane: slots[32 * k]
    enum { S_CQ = 0, S_CG, S_LQ, S_LP, S_LG, S_RQ, S_RP, S_RG, S_SQ, S_SG, S_RHO, S_CKP };
    const int S_CKR = S_CKP + (maxd + 1);
    const size_t chain = (size_t)b * a.C + c;
    const RngKey key{(uint32_t)(a.seed & 0xffffffffu), (uint32_t)(a.seed >> 32), a.chain_id0 + (uint32_t)chain};
    const int ref = a.ref[b];

    CcLane ln;
    ln.K = K;
    ln.has = lane < K;
    ln.active = ln.has && lane != ref;
    {
        const T* xu = static_cast<const T*>(a.xu) + (size_t)b * 2;
        ln.x0 = xu[0];
        ln.x1 = a.U[b] > 1 ? xu[1] : T(0);
    }
    const int jj = lane - (lane > ref);
    const int ix_b = 1 + jj, ix_i = 1 + K1 + jj, ix_a = 1 + 2 * K1 + lane;
    auto gather = [&](const T* v) {
        CcVec r;
        r.s = v[0];
        r.a = ln.has ? v[ix_a] : T(0);
        r.b = ln.active ? v[ix_b] : T(0);
        r.i = ln.active ? v[ix_i] : T(0);
        return r;
    };
    auto scatter = [&](T* v, const CcVec& r) {
        if (lane == 0) v[0] = r.s;
        if (ln.has) v[ix_a] = r.a;
        if (ln.active) {
            v[ix_b] = r.b;
            v[ix_i] = r.i;
        }
    };

    Adapt<T> ad;
    load_chain_state<T>(a, chain, vec, ad, lane, 32);
    __syncwarp();
    CcVec q = gather(vec), p = {0.f, 0.f, 0.f, 0.f}, gq = p, rho_s = p;
    T* draws = static_cast<T*>(a.draws) + chain * (size_t)a.num_results * P;
    T* stats = static_cast<T*>(a.stats) + chain * (size_t)a.num_results * SCCODA_NSTAT;

    // t = step_begin - 1: pseudo-transition that evaluates the starting point through the same (single) call site of
    // the gradient / value as the leaves of the trees.
    double lp_cur = 0.0;
#pragma unroll 1
    for (int t = a.step_begin - 1; t < a.step_end; ++t) {
        const bool boot = t < a.step_begin;
        const T eps = ad.eps;
        double e0 = 0.0;
        if (!boot) {
            T kin = T(0);
            __syncwarp();
#pragma unroll 1
            for (int i = lane; i < P; i += 32) {
                const T pm = frozen_param(a.dm, i) ? T(0) : rng_normal<T>(key, (uint32_t)i, (uint32_t)t);
                kin += pm * pm;
                vec[i] = pm;
            }
            __syncwarp();
            const CcVec pm = gather(vec), cq = cc_ld(slots + 32 * S_CQ), cg = cc_ld(slots + 32 * S_CG);
            cc_st(slots + 32 * S_LQ, cq); cc_st(slots + 32 * S_LP, pm); cc_st(slots + 32 * S_LG, cg);
            cc_st(slots + 32 * S_RQ, cq); cc_st(slots + 32 * S_RP, pm); cc_st(slots + 32 * S_RG, cg);
            cc_st(slots + 32 * S_RHO, pm);
            e0 = lp_cur - 0.5 * warp_sum_d((double)kin);
        }
        double l_lp = lp_cur, r_lp = lp_cur, c_lp = lp_cur, c_energy = e0, s_lp = lp_cur, s_en = e0;
        double wsum = 0.0, accept_sum = 0.0;
        int leapfrogs = 0, depth = 0;
        bool is_acc = false, cont = true, not_div = true;
        uint32_t leaf_serial = 0;
        const int depth_limit = boot ? 1 : maxd;
#pragma unroll 1
        while (depth < depth_limit && cont) {
            const bool forward = boot ? true : (rng_bit(key, (uint32_t)depth, (uint32_t)t, STREAM_NUTS_DIR) != 0u);
            const int e_q = forward ? S_RQ : S_LQ, e_p = forward ? S_RP : S_LP, e_g = forward ? S_RG : S_LG;
            const T se = forward ? eps : -eps;
            double lqv = forward ? r_lp : l_lp;
            if (!boot) {
                q = cc_ld(slots + 32 * e_q);
                p = cc_ld(slots + 32 * e_p);
                gq = cc_ld(slots + 32 * e_g);
                rho_s = CcVec{0.f, 0.f, 0.f, 0.f};
            }
            const unsigned n_leaves = 1u << depth;
            double w_s = -CUDART_INF;
            bool alive = true;
#pragma unroll 1
            for (unsigned leaf = 0; leaf < n_leaves && alive; ++leaf) {
                if (!boot) {
                    const T h = T(0.5) * se;
                    p.a = fmaf(h, gq.a, p.a); p.b = fmaf(h, gq.b, p.b); p.i = fmaf(h, gq.i, p.i); p.s = fmaf(h, gq.s, p.s);
                    q.a = fmaf(se, p.a, q.a); q.b = fmaf(se, p.b, q.b); q.i = fmaf(se, p.i, q.i); q.s = fmaf(se, p.s, q.s);
                }
                const CcGrad gr = cc_grad(tile, ln, q, lane);
                gq = gr.g;
                if (a.dm.freeze) gq.s = gq.i = T(0);
                lqv = cc_value(tile, pc_off, ln.has, q.a, q.b, q.i, q.s, gr.c, K, lane);
                if (boot) break;
                {
                    const T h = T(0.5) * se;
                    p.a = fmaf(h, gq.a, p.a); p.b = fmaf(h, gq.b, p.b); p.i = fmaf(h, gq.i, p.i); p.s = fmaf(h, gq.s, p.s);
                }
                rho_s.a += p.a; rho_s.b += p.b; rho_s.i += p.i; rho_s.s += p.s;
                const T kk = cc_dot(p, p, lane);
                leapfrogs += 1;
                const int idx_max = __popc(leaf >> 1);
                bool no_uturn = true;
                if ((leaf & 1u) == 0u) {
                    cc_st(slots + 32 * (S_CKP + idx_max), p);
                    cc_st(slots + 32 * (S_CKR + idx_max), rho_s);
                } else {
                    const int n_chk = __ffs(~leaf) - 1;  // trailing ones
#pragma unroll 1
                    for (int sidx = idx_max; sidx > idx_max - n_chk && no_uturn; --sidx) {
                        const CcVec cp = cc_ld(slots + 32 * (S_CKP + sidx)), cr = cc_ld(slots + 32 * (S_CKR + sidx));
                        const CcVec span = {rho_s.a - cr.a + cp.a, rho_s.b - cr.b + cp.b, rho_s.i - cr.i + cp.i, rho_s.s - cr.s + cp.s};
                        float D1, D2;
                        warp_dots2(cc_dot(span, cp, lane), cc_dot(span, p, lane), lane, D1, D2);
                        if (!(D1 >= 0.0 && D2 >= 0.0)) no_uturn = false;
                    }
                }
                double energy = lqv - 0.5 * (double)warp_sum1(kk);   // kinetic energy ~ P / 2: fp32 sum, error ~1e-6
                if (isnan(energy)) energy = -CUDART_INF;
                const double delta = energy - e0;
                const bool divergent = !(-delta < 1000.0);
                const double w_new = nuts_logaddexp<T>(w_s, delta);
                const double u = (double)rng_uniform<T>(key, leaf_serial++, (uint32_t)t, STREAM_NUTS_LEAF);
                const double thresh = (w_new > -CUDART_INF) ? delta - w_new : CUDART_NAN;
                if (nuts_log1m<T>(u) <= thresh) {
                    cc_st(slots + 32 * S_SQ, q);
                    cc_st(slots + 32 * S_SG, gq);
                    s_lp = lqv; s_en = energy;
                }
                w_s = w_new;
                accept_sum += nuts_exp_neg<T>(delta);
                if (divergent) not_div = false;
                alive = (!divergent) && no_uturn;
            }
            if (boot) {
                cc_st(slots + 32 * S_CQ, q);
                cc_st(slots + 32 * S_CG, gq);
                c_lp = lqv;
                break;
            }
            const double tree_w = alive ? w_s : -CUDART_INF;
            double thresh = tree_w - wsum;
            if (isnan(thresh)) thresh = 0.0;
            const double u = (double)rng_uniform<T>(key, (uint32_t)depth, (uint32_t)t, STREAM_NUTS_TOP);
            if ((nuts_log1m<T>(u) <= thresh) && alive) {
                cc_st(slots + 32 * S_CQ, cc_ld(slots + 32 * S_SQ));
                cc_st(slots + 32 * S_CG, cc_ld(slots + 32 * S_SG));
                c_lp = s_lp; c_energy = s_en; is_acc = true;
            }
            wsum = nuts_logaddexp<T>(wsum, tree_w);
            cc_st(slots + 32 * e_q, q);
            cc_st(slots + 32 * e_p, p);
            cc_st(slots + 32 * e_g, gq);
            CcVec rho = cc_ld(slots + 32 * S_RHO);
            rho.a += rho_s.a; rho.b += rho_s.b; rho.i += rho_s.i; rho.s += rho_s.s;
            cc_st(slots + 32 * S_RHO, rho);
            float D1, D2;
            warp_dots2(cc_dot(rho, cc_ld(slots + 32 * S_LP), lane), cc_dot(rho, cc_ld(slots + 32 * S_RP), lane), lane, D1, D2);
            if (forward) r_lp = lqv; else l_lp = lqv;
            cont = alive && (D1 >= 0.0) && (D2 >= 0.0);
            depth += 1;
        }
        lp_cur = c_lp;
        if (boot) continue;
        const double lar = accept_sum > 0.0 ? log(accept_sum / (double)leapfrogs) : -CUDART_INF;
        ad.dual(fmin(lar, 0.0), a.num_adapt, (T)a.target_accept);
        const int r = t - a.num_burnin;
        if (r >= 0 && a.store_draws) {
            scatter(draws + (size_t)r * P, cc_ld(slots + 32 * S_CQ));
            if (lane == 0)
                write_stats<T>(stats + (size_t)r * SCCODA_NSTAT, lp_cur, lar, is_acc, eps, !not_div, leapfrogs, c_energy, cont);
        }
    }
    __syncwarp();
    scatter(vec, cc_ld(slots + 32 * S_CQ));
    __syncwarp();
    store_chain_state<T>(a, chain, vec, ad, lp_cur, lane, 32);
}

#endif  // SCCODA_CC_HELPERS_ONLY

}  // namespace sccoda
