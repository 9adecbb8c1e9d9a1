// K1 — Dirichlet-multinomial spike-and-slab log-density and its analytic (digamma) gradient, evaluated
// cooperatively by one "chain group" (W warps) entirely out of shared memory.
//
// Replaces scCODAModel.target_log_prob_fn + TF autodiff:
//   /root/reference/sccoda/model/scCODA_model.py:702-760 (density), :90-100 (inputs).
//
// Layout idea ("covariate groups"): rows n with identical covariate vectors x_n have identical
// concentrations c_nk = exp(alpha_k + x_n . beta_k).  The host packer (sccoda_b200/_pack.py) sorts the rows
// of a dataset by covariate group u = grp[n] (U groups, U <= N, U <= 2^D for binary designs) so that
//   * exp / digamma(c) / lgamma(c) are evaluated U*K times instead of N*K,
//   * row sums S_u = sum_k c_uk need no cross-lane reduction over the N*K element space,
//   * only digamma(c_uk + y_nk) (and digamma(S_u + n_n)) remain per element / per row.
// With continuous covariates U == N and the scheme degenerates to the plain evaluation.
//
// State vector theta (flat, order of scCODA_model.py:692-693, 763-768):
//   [ sigma_d (D) | b_offset (D,K-1) | ind_raw (D,K-1) | alpha (K) ]
#pragma once
#include <cstdint>

#include "special.cuh"

namespace sccoda {

struct Dims {
    int N, K, D, P;
    int K1;    // K-1
    int DK1;   // D*(K-1)
    int NK;    // N*K
    uint32_t magicK, magicK1;  // ceil(2^32/K), ceil(2^32/(K-1)) for exact small-range division
};

// magic == 0 encodes a divisor of 1
__device__ __forceinline__ int fast_div(int e, uint32_t magic) {
    return magic ? (int)__umulhi((uint32_t)e, magic) : e;
}

// One dataset resident in shared memory (shared by all chain groups of the CTA).
template <typename T>
struct DataTile {
    const T* y;         // [N,K] counts after pseudocount, rows sorted by group
    const T* ycst;      // [N,K] host constant for lgamma_shift (fp32: r(y) or lgamma(y); fp64: 0)
    const T* ntot;      // [N]
    const T* xu;        // [U,D] unique covariate rows
    const int* grp;     // [N] group of row n (non-decreasing)
    const int* rstart;  // [U+1] first row of each group
    int U;
    int ref;
    double lconst;  // all theta-independent terms of the log-density (host, fp64)
};

// Scratch of one chain group.
template <typename T>
struct ChainScratch {
    T* c;        // [U,K] concentrations
    T* psic;     // [U,K] digamma(c)
    T* rowterm;  // [N]   digamma(S_u + n_n) - digamma(S_u)
    T* G;        // [N,K] d logp / d log c_nk
    T* Gu;       // [U,K] per-group column sums of G
    T* beta;     // [D,K] (reference column stays 0)
    T* ind;      // [D,K-1]
    T* gk;       // [(1+D),K]: row 0 = sum_n G_nk, row 1+d = sum_n x_nd G_nk
};

// Thread-group abstraction: W warps cooperating on one chain.  W == 1 needs no block barrier at all.
template <int W>
struct Group {
    int tid;      // thread index inside the group
    int gid;      // group slot inside the CTA (named barrier id = gid + 1)
    double* red;  // [2*W] reduction scratch (W > 1)
    static constexpr int SIZE = 32 * W;

    __device__ __forceinline__ void sync() const {
        if (W == 1) {
            __syncwarp();
        } else {
            asm volatile("bar.sync %0, %1;" ::"r"(gid + 1), "n"(32 * W) : "memory");
        }
    }
    template <typename V>
    __device__ __forceinline__ V sum(V v) const {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (W > 1) {
            V* r = reinterpret_cast<V*>(red);
            sync();  // previous consumers of red are done
            if ((tid & 31) == 0) r[tid >> 5] = v;
            sync();
            v = r[0];
#pragma unroll
            for (int w = 1; w < W; ++w) v += r[w];
        }
        return v;
    }
};

// Evaluate gradient (always) and value (WITH_LOGP) at theta.  theta/grad live in shared memory.
// On return grad[] is complete and visible to the whole group (trailing sync included).
// The value is returned to every thread (group-reduced); accumulation is fp64 in both builds.
template <typename T, bool WITH_LOGP, int W>
__device__ __forceinline__ double dm_eval(const Dims& dm, const DataTile<T>& dt, const ChainScratch<T>& cs,
                                          const T* __restrict__ theta, T* __restrict__ grad, const Group<W>& g) {
    const int tid = g.tid;
    constexpr int TS = Group<W>::SIZE;
    const int K = dm.K, D = dm.D, K1 = dm.K1, DK1 = dm.DK1, U = dt.U, ref = dt.ref;
    const T* sigma = theta;
    const T* boff = theta + D;
    const T* iraw = theta + D + DK1;
    const T* alpha = theta + D + 2 * DK1;
    double lp = 0.0;

    // ---- phase 1: spike-and-slab effects  beta = sigmoid(50 ind_raw) * sigma_d * b_offset   (:724-734)
    for (int i = tid; i < DK1; i += TS) {
        const int d = fast_div(i, dm.magicK1);
        const int j = i - d * K1;
        const int k = j + (j >= ref);
        const T bo = boff[i], ir = iraw[i];
        const T ind = sigmoid_stable(T(50) * ir);
        cs.ind[i] = ind;
        cs.beta[d * K + k] = ind * sigma[d] * bo;
        if (WITH_LOGP) lp += (double)(T(-0.5) * (bo * bo + ir * ir));  // Normal(0,1) x2   (:709-722)
    }
    if (WITH_LOGP) {
        for (int k = tid; k < K; k += TS) lp += (double)(alpha[k] * alpha[k] * T(-0.02));  // Normal(0,5) (:736-741)
        for (int d = tid; d < D; d += TS) {                                                // HalfCauchy(0,1) (:703-707)
            const T s = sigma[d];
            if (s >= T(0)) lp -= log1p((double)s * (double)s);
        }
    }
    g.sync();

    // ---- phase 2: per covariate group  c_uk = exp(alpha_k + x_u . beta_k), digamma(c_uk)          (:743)
    for (int e = tid; e < U * K; e += TS) {
        const int u = fast_div(e, dm.magicK);
        const int k = e - u * K;
        T eta = alpha[k];
        for (int d = 0; d < D; ++d) eta = fma(dt.xu[u * D + d], cs.beta[d * K + k], eta);
        const T c = exp_acc(eta);
        cs.c[e] = c;
        cs.psic[e] = digamma_pos(c);
        if (WITH_LOGP) lp -= (double)((T)(dt.rstart[u + 1] - dt.rstart[u]) * lgamma_pos(c));
    }
    g.sync();

    // ---- phase 3: rows  S_u, digamma(S_u + n_n) - digamma(S_u)
    for (int n = tid; n < dm.N; n += TS) {
        const int u = dt.grp[n];
        const T* cu = cs.c + u * K;
        T S = T(0);
        for (int k = 0; k < K; ++k) S += cu[k];
        const T nt = dt.ntot[n];
        cs.rowterm[n] = digamma_pos(S + nt) - digamma_pos(S);
        // the two O(n ln n)-sized terms are taken in fp64 (2 per row per value evaluation)
        if (WITH_LOGP) lp += lgamma((double)S) - lgamma((double)S + (double)nt);
    }
    g.sync();

    // ---- phase 4: elements  G_nk = c (digamma(c+y) - digamma(c) - rowterm_n)
    {
        const int iters = (dm.NK + TS - 1) / TS;
        for (int it = 0; it < iters; ++it) {
            const int e = tid + it * TS;
            const bool valid = e < dm.NK;
            const int ee = valid ? e : 0;
            const int n = fast_div(ee, dm.magicK);
            const int k = ee - n * K;
            const int uk = dt.grp[n] * K + k;
            const T yv = dt.y[ee];
            const T c = cs.c[uk];
            const T z = c + yv;
            // warp-uniform choice: when every lane's count is >= 6 no recurrence shift is needed
            const bool big = __all_sync(0xffffffffu, (yv >= T(6)) || !valid);
            const T psi = big ? digamma_big(z) : digamma_pos(z);
            if (valid) {
                cs.G[e] = c * (psi - cs.psic[uk] - cs.rowterm[n]);
                if (WITH_LOGP) lp += (double)lgamma_shift(c, yv, sizeof(T) == 4 ? dt.ycst[e] : T(0));
            }
        }
    }
    g.sync();

    // ---- phase 5: per-group column sums  Gu_uk = sum_{n in u} G_nk
    for (int e = tid; e < U * K; e += TS) {
        const int u = fast_div(e, dm.magicK);
        const int k = e - u * K;
        T s = T(0);
        for (int n = dt.rstart[u]; n < dt.rstart[u + 1]; ++n) s += cs.G[n * K + k];
        cs.Gu[e] = s;
    }
    g.sync();

    // ---- phase 6: gk[0,k] = sum_n G_nk ; gk[1+d,k] = sum_n x_nd G_nk
    for (int e = tid; e < (D + 1) * K; e += TS) {
        const int r = fast_div(e, dm.magicK);
        const int k = e - r * K;
        T s = T(0);
        if (r == 0) {
            for (int u = 0; u < U; ++u) s += cs.Gu[u * K + k];
        } else {
            for (int u = 0; u < U; ++u) s = fma(dt.xu[u * D + (r - 1)], cs.Gu[u * K + k], s);
        }
        cs.gk[e] = s;
    }
    g.sync();

    // ---- phase 7: chain rule back to the unconstrained parameters
    for (int i = tid; i < dm.P; i += TS) {
        T gr;
        if (i < D) {
            const int d = i;
            T acc = T(0);
            for (int j = 0; j < K1; ++j) {
                const int k = j + (j >= ref);
                acc = fma(cs.gk[(1 + d) * K + k] * cs.ind[d * K1 + j], boff[d * K1 + j], acc);
            }
            const T s = sigma[d];
            // HalfCauchy on the unconstrained sigma_d: zero prior gradient for sigma_d < 0 [TFP]
            gr = acc - (s >= T(0) ? T(2) * s / (T(1) + s * s) : T(0));
        } else if (i < D + 2 * DK1) {
            const bool is_off = i < D + DK1;
            const int ii = is_off ? i - D : i - D - DK1;
            const int d = fast_div(ii, dm.magicK1);
            const int j = ii - d * K1;
            const int k = j + (j >= ref);
            const T gg = cs.gk[(1 + d) * K + k];
            const T ind = cs.ind[ii];
            const T s = sigma[d];
            if (is_off) gr = gg * ind * s - boff[ii];
            else gr = gg * s * boff[ii] * T(50) * ind * (T(1) - ind) - iraw[ii];
        } else {
            const int k = i - D - 2 * DK1;
            gr = cs.gk[k] - alpha[k] * T(0.04);
        }
        grad[i] = gr;
    }
    g.sync();

    if (WITH_LOGP) {
        lp = g.sum(lp) + dt.lconst;
        // HalfCauchy support: -inf as soon as one sigma_d < 0 [TFP]; every thread scans the D values (uniform)
        for (int d = 0; d < D; ++d)
            if (sigma[d] < T(0)) lp = -CUDART_INF;
    }
    return lp;
}

}  // namespace sccoda
