// K1 (grouped form) — gradient of the Dirichlet-multinomial spike-and-slab log-density when the design has few
// covariate groups (case/control, factorial designs), organised by *pairs* and *items* instead of single counts.
//
// Same mathematics as dm_grad in model.cuh (reference: sccoda/model/scCODA_model.py:702-760); the data layout is
// built by sccoda_b200/_pack.py: build_items and described in include/sccoda_b200.h.
//
//   pair p = (covariate group u, cell type k), index u*K + k, concentration c_p = exp(alpha_k + x_u . beta_k);
//            plus one "total pair" per group, index Umax*K + u, concentration S_u = sum_k c_uk, counts = row totals.
//   Br_p   = sum_{n in u} [ psi(c_p + y_n) - psi(c_p) ]
//          = sum_{items of p} A_i  +  sum_{i<6} W_pi / (c_p + i)  +  sum_{odd counts} [ psi(c_p + y) - psi(c_p + 6) ]
//            (every item and the odd group of a pair write one of the pair's NF result slots)
//   A_i    = sum_{j<R} [ psi(c_p + y_ij) - psi(c_p + 6) ]          (counts >= 6, R per item, padded with 6)
//   d logp / d log c_uk = c_uk ( Br_(u,k) - Br_(total u) )
//
// Every term of Br_p is non-negative (no cancellation), psi(c) itself is never formed, and an item costs one
// lg2 of a product plus R reciprocals: the row term psi(S+n) - psi(S) of the classic formulation has become the
// bracket of the total pair and needs no per-count work at all.
#pragma once
#include "model.cuh"

namespace sccoda {

#define SCCODA_ITEM_R 5

// Per-chain record of a pair (16 bytes in the fp32 build: one LDS.128).
//   fp32: a = w6^R with w6 = 1/(c+6),  b = R * e(w6), e(w) = psi(1/w) - ln(1/w)  (what an item needs to refer its
//         log-product and its series to c+6);   fp64: a = psi(c+6), b unused.
//   rat  = sum_i W_i / (c + i)
template <typename T>
struct alignas(4 * sizeof(T)) PairRec {
    T c, a, b, rat;
    static constexpr bool has_totals = true;   // record Umax*K + u holds S_u = sum_k c_uk
};

// psi(z) - ln z = -1/(2z) - 1/(12 z^2) + 1/(120 z^4) - 1/(252 z^6) as w * s(w), w = 1/z <= 1/6; returns s(w)
__device__ __forceinline__ float psi_series_s(float w) {
    const float w2 = w * w;
    const float t = fmaf(w2, fmaf(w2, -1.0f / 252.0f, 1.0f / 120.0f), -1.0f / 12.0f);
    return fmaf(w, t, -0.5f);
}

// sum_i W_i/(c+i) from the numerators of the three quadratics c(c+5), (c+1)(c+4), (c+2)(c+3)
template <typename T>
__device__ __forceinline__ T pair_rational(T c, const T* __restrict__ co);
template <>
__device__ __forceinline__ float pair_rational<float>(float c, const float* __restrict__ co) {
    const float a = c * (c + 5.0f);
    const float r0 = rcp_fast(a), r1 = rcp_fast(a + 4.0f), r2 = rcp_fast(a + 6.0f);
    // the six coefficients of a pair are 8-byte aligned (24-byte records from a 16-byte aligned base)
    const float2* co2 = reinterpret_cast<const float2*>(co);
    const float2 n0 = co2[0], n1 = co2[1], n2 = co2[2];
    return fmaf(fmaf(n0.x, c, n0.y), r0, fmaf(fmaf(n1.x, c, n1.y), r1, fmaf(n2.x, c, n2.y) * r2));
}
template <>
__device__ __forceinline__ double pair_rational<double>(double c, const double* __restrict__ co) {
    const double a = c * (c + 5.0);
    return (co[0] * c + co[1]) / a + (co[2] * c + co[3]) / (a + 4.0) + (co[4] * c + co[5]) / (a + 6.0);
}

__device__ __forceinline__ PairRec<float> make_rec(float c, const float* __restrict__ co) {
    PairRec<float> r;
    r.c = c;
    const float w6 = rcp_fast(c + 6.0f);
    const float w2 = w6 * w6;
    r.a = w2 * w2 * w6;                                        // w6^SCCODA_ITEM_R
    r.b = (float)SCCODA_ITEM_R * (w6 * psi_series_s(w6));
    r.rat = pair_rational<float>(c, co);
    return r;
}
__device__ __forceinline__ PairRec<double> make_rec(double c, const double* __restrict__ co) {
    PairRec<double> r;
    r.c = c;
    r.a = digamma_pos(c + 6.0);
    r.b = 0.0;
    r.rat = pair_rational<double>(c, co);
    return r;
}

// A_i = sum_j [psi(c + y_j) - psi(c + 6)] for the R counts of one item (all >= 6)
__device__ __forceinline__ float item_eval(const PairRec<float>& r, const float (&y)[SCCODA_ITEM_R]) {
    static_assert(SCCODA_ITEM_R == 5, "item_eval is written for 5 counts per item");
    float z[5], acc = -r.b;
#pragma unroll
    for (int j = 0; j < 5; ++j) z[j] = r.c + y[j];
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const float w = rcp_fast(z[j]);
        acc = fmaf(w, psi_series_s(w), acc);
    }
    // prod_j z_j / (c+6)^5 >= 1: the logarithm is taken of the ratio, so its rounding error scales with the result
    const float prod = ((z[0] * z[1]) * r.a) * ((z[2] * z[3]) * z[4]);
    return fmaf(lg2_fast(prod), SCCODA_LN2F, acc);
}
__device__ __forceinline__ double item_eval(const PairRec<double>& r, const double (&y)[SCCODA_ITEM_R]) {
    double acc = 0.0;
#pragma unroll 1
    for (int j = 0; j < SCCODA_ITEM_R; ++j) acc += digamma_pos(r.c + y[j]) - r.a;
    return acc;
}

// psi(c + y) - psi(c + 6) for a count that is neither >= 6 nor in {0,...,5} (rare)
__device__ __forceinline__ float odd_eval(const PairRec<float>& r, float y) {
    return digamma_pos(r.c + y) - digamma_big(r.c + 6.0f);
}
__device__ __forceinline__ double odd_eval(const PairRec<double>& r, double y) { return digamma_pos(r.c + y) - r.a; }

// Br_p = rational + the NF result slots of the pair (items and odd group; unused slots hold 0)
template <typename T>
__device__ __forceinline__ T pair_bracket(T rat, const T* __restrict__ slots, int NF) {
    T s0 = rat, s1 = T(0);
#pragma unroll 1
    for (int t = 0; t < NF; t += 2) {
        s0 += slots[t];
        s1 += slots[t + 1];
    }
    return s0 + s1;
}
template <>
__device__ __forceinline__ float pair_bracket<float>(float rat, const float* __restrict__ slots, int NF) {
    if (NF == 4) {   // the usual case: two items and an odd group at most
        const float4 v = *reinterpret_cast<const float4*>(slots);
        return (rat + v.x) + (v.y + (v.z + v.w));
    }
    if (NF == 2) {
        const float2 v = *reinterpret_cast<const float2*>(slots);
        return (rat + v.x) + v.y;
    }
    float s0 = rat, s1 = 0.0f;
#pragma unroll 1
    for (int t = 0; t < NF; t += 2) {
        const float2 v = *reinterpret_cast<const float2*>(slots + t);
        s0 += v.x;
        s1 += v.y;
    }
    return s0 + s1;
}

// Gradient of the log-density at theta, grouped form.  Same contract as dm_grad (model.cuh): every component is
// handed once to epi(i, value) by its owner, a trailing group sync publishes what the epilogue wrote, and the
// scratch keeps the concentrations (PairRec::c, pairs u*K + k and Umax*K + u) for dm_value.
//
// The pass comes in two halves so that a caller can run the second half and the value pass side by side on two halves
// of the chain's warps (they only share read access to the pair records): `front` = effects and pair records (P0, P1),
// `back` = items, brackets and the chain rule (E, P2, C).
template <typename T, int W>
__device__ __forceinline__ void dm_grad_items_front(const Dims& dm, const DataTile<T>& dt, const ChainScratch& cs,
                                                    uint32_t theta_off, const Group<W>& g) {
    const int tid = g.tid;
    constexpr int TS = Group<W>::SIZE;
    const int K = dm.K, D = dm.D, K1 = dm.K1, DK1 = dm.DK1, U = dt.U, ref = dt.ref, tot0 = dt.tot0;
    const T* sigma = sp<T>(theta_off);
    const T* boff = sigma + D;
    const T* iraw = boff + DK1;
    const T* alpha = iraw + DK1;
    PairRec<T>* prec = sp<PairRec<T>>(cs.cp);
    T* Ssm = sp<T>(cs.Gu);
    T* beta = sp<T>(cs.beta);
    T* ind = sp<T>(cs.ind);
    const T* xu = sp<T>(dt.xu);
    const T* pcoef = sp<T>(dt.pcoef);
    const int lane = tid & 31, warp = tid >> 5;

    // ---- P0: spike-and-slab effects, one column per thread                                        (:724-734)
#pragma unroll 1
    for (int k = tid; k < K; k += TS) {
        if (k != ref) {
            const int j = k - (k > ref);
#pragma unroll 1
            for (int d = 0; d < D; ++d) {
                const int i = d * K1 + j;
                const T id = sigmoid_stable(T(50) * iraw[i]);
                ind[i] = id;
                beta[d * K + k] = id * sigma[d] * boff[i];
            }
        }
    }
    // ---- P1: pair records c_uk = exp(alpha_k + x_u . beta_k) (:743) and the total pairs S_u = sum_k c_uk
    if (W == 1 && U <= 32 && K <= 64) {
        // one warp, few groups: group-major passes (thread k reads the beta it wrote itself: no sync), S_u by shuffles
        T mine = T(1);
#pragma unroll 1
        for (int u = 0; u < U; ++u) {
            T part = T(0);
#pragma unroll 1
            for (int k = lane; k < K; k += 32) {
                T eta = alpha[k];
#pragma unroll 1
                for (int d = 0; d < D; ++d) eta = fma(xu[u * D + d], beta[d * K + k], eta);
                const T c = exp_acc(eta);
                prec[u * K + k] = make_rec(c, pcoef + 6 * (u * K + k));
                part += c;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
            if (lane == u) mine = part;
        }
        if (lane < U) prec[tot0 + lane] = make_rec(mine, pcoef + 6 * (tot0 + lane));
    } else {
        g.sync();
#pragma unroll 1
        for (int p = tid; p < U * K; p += TS) {
            const int u = fast_div(p, dm.magicK);
            const int k = p - u * K;
            T eta = alpha[k];
#pragma unroll 1
            for (int d = 0; d < D; ++d) eta = fma(xu[u * D + d], beta[d * K + k], eta);
            prec[p] = make_rec(exp_acc(eta), pcoef + 6 * p);
        }
        g.sync();
#pragma unroll 1
        for (int u = warp; u < U; u += W) {   // one warp per group
            T part = T(0);
#pragma unroll 1
            for (int k = lane; k < K; k += 32) part += prec[u * K + k].c;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
            if (lane == 0) Ssm[u] = part;
        }
        g.sync();
#pragma unroll 1
        for (int u = tid; u < U; u += TS) prec[tot0 + u] = make_rec(Ssm[u], pcoef + 6 * (tot0 + u));
    }
    g.sync();
}

template <typename T, int W, class Epi>
__device__ __forceinline__ void dm_grad_items_back(const Dims& dm, const DataTile<T>& dt, const ChainScratch& cs,
                                                   uint32_t theta_off, const Group<W>& g, const Epi& epi) {
    const int tid = g.tid;
    constexpr int TS = Group<W>::SIZE;
    constexpr int R = SCCODA_ITEM_R;
    const int K = dm.K, D = dm.D, K1 = dm.K1, DK1 = dm.DK1, U = dt.U, ref = dt.ref, tot0 = dt.tot0, NF = dt.NF;
    const T* sigma = sp<T>(theta_off);
    const T* boff = sigma + D;
    const T* iraw = boff + DK1;
    const T* alpha = iraw + DK1;
    PairRec<T>* prec = sp<PairRec<T>>(cs.cp);
    T* A = sp<T>(cs.G);          // result slots [NP*NF + 1]
    const T* ind = sp<T>(cs.ind);
    const T* xu = sp<T>(dt.xu);
    const T* iy = sp<T>(dt.iy);
    const uint16_t* ipair = sp<uint16_t>(dt.ipair);
    const uint16_t* ipos = sp<uint16_t>(dt.ipos);
    // ---- E: items (one per thread and pass) and odd groups (rare), each into its own result slot
    {
        const int NIs = dt.NImax;
#pragma unroll 1
        for (int i = tid; i < dt.NI; i += TS) {
            const PairRec<T> r = prec[ipair[i]];
            T y[R];
#pragma unroll
            for (int j = 0; j < R; ++j) y[j] = iy[j * NIs + i];
            A[ipos[i]] = item_eval(r, y);
        }
#pragma unroll 1
        for (int q = tid; q < dt.NG; q += TS) {
            const PairRec<T> r = prec[dt.opair[q]];
            const uint32_t od = dt.odesc[q];
            const T* oy = dt.oy + (od & 0xffffu);
            T s = T(0);
#pragma unroll 1
            for (int t = 0; t < (int)(od >> 16); ++t) s += odd_eval(r, oy[t]);
            A[dt.opos[q]] = s;
        }
    }
    g.sync();
    // ---- P2: one pair per thread: G_uk = c_uk (Br_uk - Br_total(u)), parked in the (no longer needed) rational slot
#pragma unroll 1
    for (int p = tid; p < U * K; p += TS) {
        const int pt = tot0 + fast_div(p, dm.magicK);
        const T bt = pair_bracket<T>(prec[pt].rat, A + pt * NF, NF);
        const PairRec<T> r = prec[p];
        prec[p].rat = r.c * (pair_bracket<T>(r.rat, A + p * NF, NF) - bt);
    }
    g.sync();
    // ---- C: one column per thread: column sums over the groups, chain rule, sigma_d by a group
    //         reduction.  Covariates are taken four at a time.
#pragma unroll 1
    for (int d0 = 0; d0 < D; d0 += 4) {
        T part[4] = {T(0), T(0), T(0), T(0)};
#pragma unroll 1
        for (int k0 = 0; k0 < K; k0 += TS) {
            const int k = k0 + tid;
            const bool has = k < K;
            const int kc = has ? k : 0;
            T g0 = T(0);
            T acc[4] = {T(0), T(0), T(0), T(0)};
#pragma unroll 1
            for (int u = 0; u < U; ++u) {
                const T G = prec[u * K + kc].rat;
                g0 += G;
#pragma unroll
                for (int dd = 0; dd < 4; ++dd)
                    if (d0 + dd < D) acc[dd] = fma(xu[u * D + d0 + dd], G, acc[dd]);
            }
            if (has) {
                if (d0 == 0) epi(D + 2 * DK1 + k, fma(alpha[k], T(-0.04), g0));          // d/d alpha_k
                if (k != ref) {
                    const int j = k - (k > ref);
#pragma unroll
                    for (int dd = 0; dd < 4; ++dd) {
                        if (d0 + dd < D) {
                            const int d = d0 + dd;
                            const int i = d * K1 + j;
                            const T s = sigma[d];
                            const T id = ind[i];
                            const T bo = boff[i];
                            const T t = acc[dd] * id;
                            epi(D + i, fma(t, s, -bo));                                          // d/d b_offset_dk
                            epi(D + DK1 + i, fma(t * s * bo, T(50) * (T(1) - id), -iraw[i]));    // d/d ind_raw_dk
                            part[dd] = fma(t, bo, part[dd]);
                        }
                    }
                }
            }
        }
        // d/d sigma_d: sum_k g_dk ind_dk b_offset_dk - HalfCauchy term (zero prior gradient below 0 [TFP]); the
        // (up to) four sums of this block of covariates share one reduction
        if constexpr (sizeof(T) == 4) {
            g.sumv(part);
        } else {
#pragma unroll
            for (int dd = 0; dd < 4; ++dd)
                if (d0 + dd < D) part[dd] = g.sum(part[dd]);
        }
#pragma unroll
        for (int dd = 0; dd < 4; ++dd) {
            if (d0 + dd < D) {
                const T s = sigma[d0 + dd];
                if (tid == 0) epi(d0 + dd, part[dd] - (s >= T(0) ? T(2) * s / (T(1) + s * s) : T(0)));
            }
        }
    }
    g.sync();
}

template <typename T, int W, class Epi>
__device__ __forceinline__ void dm_grad_items(const Dims& dm, const DataTile<T>& dt, const ChainScratch& cs,
                                              uint32_t theta_off, const Group<W>& g, const Epi& epi) {
    dm_grad_items_front<T, W>(dm, dt, cs, theta_off, g);
    dm_grad_items_back<T, W>(dm, dt, cs, theta_off, g, epi);
}

}  // namespace sccoda
