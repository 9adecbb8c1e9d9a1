// K1 — Dirichlet-multinomial spike-and-slab log-density and its analytic (digamma) gradient, evaluated
// cooperatively by one "chain group" (W warps) out of shared memory.
//
// Replaces scCODAModel.target_log_prob_fn + TF autodiff:
//   /root/reference/sccoda/model/scCODA_model.py:702-760 (density), :90-100 (inputs).
//
// Data layout (built by sccoda_b200/_pack.py, see include/sccoda_b200.h):
//  * "covariate groups": rows n with identical covariate vectors x_n share their concentrations
//    c_nk = exp(alpha_k + x_n . beta_k).  Rows are sorted by group u = grp[n] (U groups; U <= 2^D for binary
//    designs, U == N for continuous covariates), so exp / digamma(c) are evaluated U*K times, not N*K, and the
//    row sums S_u need no reduction over the N*K element space.
//  * "element list": the N*K counts are stored column-major with the columns (cell types) sorted by decreasing
//    minimum count.  Element e = kk*N + n carries y_e, a host constant ycst_e and a packed index
//    (u*K + k) | (n << 16).  The first `nbig` elements belong to columns whose counts are all >= 6: for them
//    digamma(c + y) needs no recurrence shift (2 MUFU + 6 FMA, branch-free, no warp vote); the few remaining
//    elements take the general digamma.  (y, index) pairs are staged once into shared memory as 8-byte
//    records (one LDS.64 per element), (c, digamma(c)) pairs likewise.
//
// All shared-memory arrays are addressed as 32-bit byte offsets from the dynamic shared-memory base, so that
// every access is an LDS/STS with a 32-bit address (no generic-pointer arithmetic, fewer registers).
//
// State vector theta (flat, order of scCODA_model.py:692-693, 763-768):
//   [ sigma_d (D) | b_offset (D,K-1) | ind_raw (D,K-1) | alpha (K) ]
#pragma once
#include <cstdint>

#include "special.cuh"

namespace sccoda {

extern __shared__ __align__(16) unsigned char sccoda_smem[];

// typed view of a shared-memory array given its byte offset
template <typename V>
__device__ __forceinline__ V* sp(uint32_t off) {
    return reinterpret_cast<V*>(sccoda_smem + off);
}

struct Dims {
    int N, K, D, P;
    int K1;    // K-1
    int DK1;   // D*(K-1)
    int NK;    // N*K
    uint32_t magicK, magicK1;  // ceil(2^32/K), ceil(2^32/(K-1)); 0 encodes a divisor of 1
    int freeze;                // 1: sigma_d and ind_raw are frozen (zero momentum, zero gradient): SimpleModel
    double conc_max;           // states with a concentration above this get a NaN value => rejected (sccoda_sampler::reject_conc_above)
};

// is parameter i one of the frozen ones (sigma_d, ind_raw) of a run with Dims::freeze?
__device__ __forceinline__ bool frozen_param(const Dims& dm, int i) {
    return dm.freeze && (i < dm.D || (i >= dm.D + dm.DK1 && i < dm.D + 2 * dm.DK1));
}

__device__ __forceinline__ int fast_div(int e, uint32_t magic) {
    return magic ? (int)__umulhi((uint32_t)e, magic) : e;
}

// One element of the list: count and packed index (u*K + k) | (n << 16); 8 bytes in the fp32 build (LDS.64).
template <typename T>
struct alignas(2 * sizeof(T)) Elem {
    T y;
    uint32_t idx;
};
// concentration and its digamma, read together
template <typename T>
struct alignas(2 * sizeof(T)) CPsi {
    T c;
    T psi;
    static constexpr bool has_totals = false;
};

// One dataset (shared by all chain groups of the CTA): shared-memory offsets + global pointers.
template <typename T>
struct DataTile {
    uint32_t ntot;         // smem T[N]
    uint32_t nrc;          // smem T[N] grouped fp32 build: r(n) (n >= 6) or lgamma(n) of the row totals, cf. eycst
    uint32_t xu;           // smem T[U,D] unique covariate rows
    uint32_t grp;          // smem int[N] group of row n (non-decreasing)
    uint32_t rstart;       // smem int[U+1] first row of each group
    uint32_t colperm;      // smem int[K] sorted column position -> original cell type
    uint32_t el;           // smem Elem<T>[NK] element list (only when el_smem)
    int el_smem;
    const T* ey;           // global [NK]
    const uint32_t* eidx;  // global [NK]
    const T* eycst;        // global [NK] host constants of the value pass
    uint32_t eyc;          // smem T[NK] copy of eycst (only when el_smem)
    int U, ref;
    int nbig;              // elements [0, nbig) have counts >= 6
    double lconst;         // all theta-independent terms of the log-density (host, fp64)
    // pair / item layout (grouped gradient path, grouped.cuh)
    uint32_t pcoef;        // smem T[NP,6] rational coefficients of each pair
    uint32_t iy;           // smem T[R,NImax] counts of the items
    uint32_t ipair;        // smem uint16[NImax] pair of each item
    uint32_t ipos;         // smem uint16[NImax] result slot of each item
    uint32_t inreal;       // smem uint8[NImax] data counts of each item (value pass)
    const T* oy;           // global [NOmax] odd counts
    const uint16_t* opair; // global [NGmax] pair / result slot / (first | count << 16) of each odd group
    const uint16_t* opos;
    const uint32_t* odesc;
    int NI, NG, NImax, NF; // items / odd groups of this dataset, item stride, result slots per pair
    int tot0;              // index of the first total pair (Umax*K)
};

// Scratch of one chain group (shared-memory byte offsets).
struct ChainScratch {
    uint32_t cp;       // CPsi<T>[U,K] concentrations c_uk and digamma(c_uk)   | grouped: PairRec<T>[Umax*(K+1)]
    uint32_t rowterm;  // T[N]   digamma(S_u + n_n) - digamma(S_u)              | grouped: unused
    uint32_t G;        // T[NK]  d logp / d log c, element-list order           | grouped: T[NP*NF+1] result slots
    uint32_t Gu;       // T[U]   row sums S_u of the concentrations (small-U path)
    uint32_t beta;     // T[D,K] (reference column stays 0)
    uint32_t ind;      // T[D,K-1]
    uint32_t gk;       // T[(1+D),K]: row 0 = sum_n G_nk, row 1+d = sum_n x_nd G_nk
    uint32_t vec;      // T[nvec, vs] state vectors of the sampler
};

// Thread-group abstraction: W warps cooperating on one chain.  W == 1 needs no block barrier at all.
template <int W>
struct Group {
    int tid;       // thread index inside the group
    int gid;       // group slot inside the CTA (named barrier id = gid + 1)
    uint32_t red;  // smem double[2*W] reduction scratch (W > 1)
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
            V* r = sp<V>(red);
            sync();  // previous consumers of red are done
            if ((tid & 31) == 0) r[tid >> 5] = v;
            sync();
            // second stage as a tree over the W warp sums (a serial walk costs W dependent shared-memory loads)
            if constexpr ((W & (W - 1)) == 0) {
                v = r[tid & (W - 1)];
#pragma unroll
                for (int o = W >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            } else {   // any number of warps (the split groups of grad_value_of): lanes >= W carry zeros
                v = (tid & 31) < W ? r[tid & 31] : V(0);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            }
        }
        return v;
    }
    // NV sums at once with a single pair of barriers (the scratch holds 16 bytes per warp: two doubles / four floats)
    template <typename V, int NV>
    __device__ __forceinline__ void sumv(V (&v)[NV]) const {
        static_assert(sizeof(V) * NV <= 16, "Group::sumv: 16 bytes of scratch per warp");
#pragma unroll
        for (int j = 0; j < NV; ++j)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v[j] += __shfl_xor_sync(0xffffffffu, v[j], o);
        if (W > 1) {
            V* r = sp<V>(red);
            sync();
            if ((tid & 31) == 0)
#pragma unroll
                for (int j = 0; j < NV; ++j) r[(tid >> 5) * NV + j] = v[j];
            sync();
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                V t;
                if constexpr ((W & (W - 1)) == 0) {
                    t = r[(tid & (W - 1)) * NV + j];
#pragma unroll
                    for (int o = W >> 1; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
                } else {
                    t = (tid & 31) < W ? r[(tid & 31) * NV + j] : V(0);
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
                }
                v[j] = t;
            }
        }
    }
    // group-wide OR of a predicate
    __device__ __forceinline__ bool any(bool v) const {
        v = __any_sync(0xffffffffu, v) != 0;
        if (W > 1) {
            int* r = sp<int>(red);
            sync();
            if ((tid & 31) == 0) r[tid >> 5] = v ? 1 : 0;
            sync();
            int o = 0;
#pragma unroll
            for (int w = 0; w < W; ++w) o |= r[w];
            v = o != 0;
        }
        return v;
    }
};

// Epilogue that only stores the gradient.
template <typename T>
struct StoreGrad {
    T* grad;
    __device__ __forceinline__ void operator()(int i, T gr) const { grad[i] = gr; }
};

// Gradient of the log-density at theta (shared-memory offset of T[P]).  Every component i of the gradient is
// handed exactly once to epi(i, value) by the thread that owns it (the sampler fuses its leapfrog update there;
// the epilogue may read theta[i] but must not write the theta buffer).  A trailing group sync makes everything
// the epilogue wrote visible.  The scratch keeps c_uk for dm_value.
//
//   F1  columns k:   ind_dk = sigmoid(50 ind_raw), beta_dk = ind sigma_d b_offset              (:724-734)
//   F2  (u, k):      c_uk = exp(alpha_k + x_u . beta_k), digamma(c_uk), S_u = sum_k c_uk         (:743)
//   F3  rows n:      rowterm_n = digamma(S_u + n_n) - digamma(S_u)
//   E   elements:    G_e = c (digamma(c + y) - digamma(c) - rowterm_n)
//   B1  (r, k):      gk[r,k] = sum_n w_r(n) G_nk,  w_0 = 1, w_{1+d}(n) = x_nd
//   B2  columns k:   chain rule to (alpha_k, b_offset_dk, ind_raw_dk), sigma_d by a group reduction
// The pass comes in two halves (cf. dm_grad_items): `front` = effects, concentrations and row terms (F1-F3), `back` =
// elements, column sums and the chain rule (E, B); the value pass only reads what the front half wrote.
template <typename T, int W>
__device__ __forceinline__ void dm_grad_front(const Dims& dm, const DataTile<T>& dt, const ChainScratch& cs,
                                              uint32_t theta_off, const Group<W>& g) {
    const int tid = g.tid;
    constexpr int TS = Group<W>::SIZE;
    const int K = dm.K, D = dm.D, K1 = dm.K1, DK1 = dm.DK1, N = dm.N, U = dt.U, ref = dt.ref;
    const T* sigma = sp<T>(theta_off);
    const T* boff = sigma + D;
    const T* iraw = boff + DK1;
    const T* alpha = iraw + DK1;
    CPsi<T>* cp = sp<CPsi<T>>(cs.cp);
    T* rowterm = sp<T>(cs.rowterm);
    T* Ssm = sp<T>(cs.Gu);  // [U] row sums of the concentrations (small-U path)
    T* beta = sp<T>(cs.beta);
    T* ind = sp<T>(cs.ind);
    const T* xu = sp<T>(dt.xu);
    const T* ntot = sp<T>(dt.ntot);
    const int* grp = sp<int>(dt.grp);
    const int* rstart = sp<int>(dt.rstart);

    // ---- F1: spike-and-slab effects of the thread's own columns (read back by the same thread in F2: no sync)
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
    // ---- F2 + F3
    if (W == 1 && U <= 32) {
        // few covariate groups, one warp: S_u by warp shuffles while the concentrations are produced
#pragma unroll 1
        for (int u = 0; u < U; ++u) {
            T part = T(0);
#pragma unroll 1
            for (int k = tid; k < K; k += TS) {
                T eta = alpha[k];
#pragma unroll 1
                for (int d = 0; d < D; ++d) eta = fma(xu[u * D + d], beta[d * K + k], eta);
                CPsi<T> v;
                v.c = exp_acc(eta);
                v.psi = digamma_pos(v.c);
                cp[u * K + k] = v;
                part += v.c;
            }
            part = g.sum(part);
            if (tid == 0) Ssm[u] = part;
        }
        g.sync();
#pragma unroll 1
        for (int n = tid; n < N; n += TS) {
            const T S = Ssm[grp[n]];
            rowterm[n] = digamma_pos(S + ntot[n]) - digamma_pos(S);
        }
    } else {
        if constexpr (W > 1) {
            // several warps per chain: one (group, column) pair per thread and pass.  (The column-owner walk below
            // leaves all but K threads idle for U dependent exp / digamma evaluations: with a continuous covariate,
            // U = N, that was 43 % of the stall samples of a NUTS leaf.)
            g.sync();
#pragma unroll 1
            for (int p = tid; p < U * K; p += TS) {
                const int u = fast_div(p, dm.magicK);
                const int k = p - u * K;
                T eta = alpha[k];
#pragma unroll 1
                for (int d = 0; d < D; ++d) eta = fma(xu[u * D + d], beta[d * K + k], eta);
                CPsi<T> v;
                v.c = exp_acc(eta);
                v.psi = digamma_pos(v.c);
                cp[p] = v;
            }
        } else {
#pragma unroll 1
            for (int k = tid; k < K; k += TS) {
                const T a = alpha[k];
#pragma unroll 1
                for (int u = 0; u < U; ++u) {
                    T eta = a;
#pragma unroll 1
                    for (int d = 0; d < D; ++d) eta = fma(xu[u * D + d], beta[d * K + k], eta);
                    CPsi<T> v;
                    v.c = exp_acc(eta);
                    v.psi = digamma_pos(v.c);
                    cp[u * K + k] = v;
                }
            }
        }
        g.sync();
#pragma unroll 1
        for (int n = tid; n < N; n += TS) {
            const CPsi<T>* cu = cp + grp[n] * K;
            T S0 = T(0), S1 = T(0);
            int k = 0;
#pragma unroll 2
            for (; k + 1 < K; k += 2) {
                S0 += cu[k].c;
                S1 += cu[k + 1].c;
            }
            if (k < K) S0 += cu[k].c;
            const T S = S0 + S1;
            rowterm[n] = digamma_pos(S + ntot[n]) - digamma_pos(S);
        }
    }
    g.sync();
}

template <typename T, int W, class Epi>
__device__ __forceinline__ void dm_grad_back(const Dims& dm, const DataTile<T>& dt, const ChainScratch& cs,
                                             uint32_t theta_off, const Group<W>& g, const Epi& epi) {
    const int tid = g.tid;
    constexpr int TS = Group<W>::SIZE;
    const int K = dm.K, D = dm.D, K1 = dm.K1, DK1 = dm.DK1, N = dm.N, U = dt.U, ref = dt.ref;
    const T* sigma = sp<T>(theta_off);
    const T* boff = sigma + D;
    const T* iraw = boff + DK1;
    const T* alpha = iraw + DK1;
    const CPsi<T>* cp = sp<CPsi<T>>(cs.cp);
    const T* rowterm = sp<T>(cs.rowterm);
    T* G = sp<T>(cs.G);
    const T* ind = sp<T>(cs.ind);
    T* gk = sp<T>(cs.gk);
    const T* xu = sp<T>(dt.xu);
    const int* grp = sp<int>(dt.grp);
    const int* rstart = sp<int>(dt.rstart);
    const int* colperm = sp<int>(dt.colperm);

    // ---- E: elements  G = c (digamma(c+y) - digamma(c) - rowterm_n)
    if (dt.el_smem) {
        const Elem<T>* el = sp<Elem<T>>(dt.el);
        // counts >= 6, no recurrence shift needed.  Full batches of EB elements per thread are written
        // loads-first / stores-last: the arrays share one shared-memory buffer, so the compiler would otherwise
        // keep every element's loads behind the previous element's store (no ILP inside the warp).
        constexpr int EB = 4;
        int e0 = tid;
#pragma unroll 1
        for (; e0 + (EB - 1) * TS < dt.nbig; e0 += EB * TS) {
            Elem<T> v[EB];
            CPsi<T> q[EB];
            T rt[EB];
#pragma unroll
            for (int j = 0; j < EB; ++j) v[j] = el[e0 + j * TS];
#pragma unroll
            for (int j = 0; j < EB; ++j) {
                q[j] = cp[v[j].idx & 0xffffu];
                rt[j] = rowterm[v[j].idx >> 16];
            }
#pragma unroll
            for (int j = 0; j < EB; ++j) rt[j] = q[j].c * (digamma_big(q[j].c + v[j].y) - q[j].psi - rt[j]);
#pragma unroll
            for (int j = 0; j < EB; ++j) G[e0 + j * TS] = rt[j];
        }
#pragma unroll 1
        for (int e = e0; e < dt.nbig; e += TS) {
            const Elem<T> v = el[e];
            const CPsi<T> q = cp[v.idx & 0xffffu];
            G[e] = q.c * (digamma_big(q.c + v.y) - q.psi - rowterm[v.idx >> 16]);
        }
        // columns that contain small counts: general digamma
#pragma unroll 1
        for (int e = dt.nbig + tid; e < dm.NK; e += TS) {
            const Elem<T> v = el[e];
            const CPsi<T> q = cp[v.idx & 0xffffu];
            G[e] = q.c * (digamma_pos(q.c + v.y) - q.psi - rowterm[v.idx >> 16]);
        }
    } else {
        // element list streamed from global memory (fp64 parity build, or tiles too large for shared memory)
#pragma unroll 1
        for (int e = tid; e < dm.NK; e += TS) {
            const uint32_t ix = dt.eidx[e];
            const CPsi<T> q = cp[ix & 0xffffu];
            G[e] = q.c * (digamma_pos(q.c + dt.ey[e]) - q.psi - rowterm[ix >> 16]);
        }
    }
    g.sync();

    // ---- B: column sums gk[0,k] = sum_n G_nk, gk[1+d,k] = sum_n x_nd G_nk and the chain rule
    if (U <= 8 && K <= TS && D <= 4) {
        // Register-resident fast path (few covariate groups, one column per thread): thread kk sums its whole
        // column (rows are contiguous per group), keeps the (1+D) sums in registers and goes straight on to the
        // parameters of that column; only the sigma_d reduction crosses threads.
        const bool has = tid < K;
        const int k = colperm[has ? tid : 0];
        T g0 = T(0);
        T acc[4] = {T(0), T(0), T(0), T(0)};
        if (has) {
            const T* p = G + tid * N;
            int n = 0;
#pragma unroll 1
            for (int u = 0; u < U; ++u) {
                const int n1 = rstart[u + 1];
                const int cnt = n1 - n;
                n = n1;
                T s0 = T(0), s1 = T(0), s2 = T(0), s3 = T(0);
#pragma unroll 1
                for (int q = cnt >> 2; q > 0; --q, p += 4) {
                    s0 += p[0];
                    s1 += p[1];
                    s2 += p[2];
                    s3 += p[3];
                }
                if (cnt & 2) {
                    s0 += p[0];
                    s1 += p[1];
                    p += 2;
                }
                if (cnt & 1) {
                    s2 += p[0];
                    p += 1;
                }
                const T su = (s0 + s1) + (s2 + s3);
                g0 += su;
#pragma unroll
                for (int d = 0; d < 4; ++d)
                    if (d < D) acc[d] = fma(xu[u * D + d], su, acc[d]);
            }
        }
        const bool active = has && (k != ref);
        const int j = k - (k > ref);
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            if (d < D) {
                const T s = sigma[d];
                T part = T(0);
                if (active) {
                    const int i = d * K1 + j;
                    const T id = ind[i];
                    const T bo = boff[i];
                    const T t = acc[d] * id;
                    epi(D + i, fma(t, s, -bo));                                          // d/d b_offset_dk
                    epi(D + DK1 + i, fma(t * s * bo, T(50) * (T(1) - id), -iraw[i]));    // d/d ind_raw_dk
                    part = t * bo;
                }
                // d/d sigma_d: sum_k g_dk ind_dk b_offset_dk - HalfCauchy term (zero prior gradient below 0 [TFP])
                part = g.sum(part);
                if (tid == 0) epi(d, part - (s >= T(0) ? T(2) * s / (T(1) + s * s) : T(0)));
            }
        }
        if (has) epi(D + 2 * DK1 + k, fma(alpha[k], T(-0.04), g0));                     // d/d alpha_k
        g.sync();
        return;
    }
    if (U <= 8) {
        // few covariate groups, several columns per thread: same walk, sums parked in shared memory
#pragma unroll 1
        for (int kk = tid; kk < K; kk += TS) {
            const int k = colperm[kk];
            const T* col = G + kk * N;
            T g0 = T(0);
            int n = 0;
#pragma unroll 1
            for (int u = 0; u < U; ++u) {
                const int n1 = rstart[u + 1];
                T s0 = T(0), s1 = T(0);
#pragma unroll 2
                for (; n + 1 < n1; n += 2) {
                    s0 += col[n];
                    s1 += col[n + 1];
                }
                if (n < n1) s0 += col[n++];
                const T su = s0 + s1;
                g0 += su;
#pragma unroll 1
                for (int d = 0; d < D; ++d) {
                    const T prev = (u == 0) ? T(0) : gk[(1 + d) * K + k];
                    gk[(1 + d) * K + k] = fma(xu[u * D + d], su, prev);
                }
            }
            gk[k] = g0;
        }
    } else if (W > 1) {
        // many groups, several warps per chain: one warp per (sum, column), the rows over its lanes
        const int lane = tid & 31;
#pragma unroll 1
        for (int e = tid >> 5; e < (D + 1) * K; e += W) {
            const int r = fast_div(e, dm.magicK);
            const int kk = e - r * K;
            const T* col = G + kk * N;
            T s = T(0);
#pragma unroll 1
            for (int n = lane; n < N; n += 32) s = fma((r == 0) ? T(1) : xu[grp[n] * D + (r - 1)], col[n], s);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) gk[r * K + colperm[kk]] = s;
        }
        g.sync();
    } else {
#pragma unroll 1
        for (int e = tid; e < (D + 1) * K; e += TS) {
            const int r = fast_div(e, dm.magicK);
            const int kk = e - r * K;
            const T* col = G + kk * N;
            T s = T(0);
            int n = 0;
#pragma unroll 1
            for (int u = 0; u < U; ++u) {
                const int n1 = rstart[u + 1];
                T s0 = T(0), s1 = T(0);
#pragma unroll 2
                for (; n + 1 < n1; n += 2) {
                    s0 += col[n];
                    s1 += col[n + 1];
                }
                if (n < n1) s0 += col[n++];
                const T w = (r == 0) ? T(1) : xu[u * D + (r - 1)];
                s = fma(w, s0 + s1, s);
            }
            gk[r * K + colperm[kk]] = s;
        }
        g.sync();
    }
    // Note on ownership: in the small-U path thread t produced gk[., colperm[kk]] for kk = t, t+TS, ...; the chain
    // rule below is walked by sorted position too (k = colperm[kk]), so every thread consumes only its own sums.
#pragma unroll 1
    for (int d = 0; d < D; ++d) {
        const T s = sigma[d];
        T part = T(0);
#pragma unroll 1
        for (int kk = tid; kk < K; kk += TS) {
            const int k = colperm[kk];
            if (k != ref) {
                const int i = d * K1 + k - (k > ref);
                const T gg = gk[(1 + d) * K + k];
                const T id = ind[i];
                const T bo = boff[i];
                const T t = gg * id;
                epi(D + i, fma(t, s, -bo));                                              // d/d b_offset_dk
                epi(D + DK1 + i, fma(t * s * bo, T(50) * (T(1) - id), -iraw[i]));        // d/d ind_raw_dk
                part = fma(t, bo, part);
            }
        }
        // d/d sigma_d: sum_k g_dk ind_dk b_offset_dk - HalfCauchy term (zero prior gradient for sigma_d < 0 [TFP])
        part = g.sum(part);
        if (tid == 0) epi(d, part - (s >= T(0) ? T(2) * s / (T(1) + s * s) : T(0)));
    }
#pragma unroll 1
    for (int kk = tid; kk < K; kk += TS) {
        const int k = colperm[kk];
        epi(D + 2 * DK1 + k, fma(alpha[k], T(-0.04), gk[k]));                            // d/d alpha_k
    }
    g.sync();
}

template <typename T, int W, class Epi>
__device__ __forceinline__ void dm_grad(const Dims& dm, const DataTile<T>& dt, const ChainScratch& cs,
                                        uint32_t theta_off, const Group<W>& g, const Epi& epi) {
    dm_grad_front<T, W>(dm, dt, cs, theta_off, g);
    dm_grad_back<T, W>(dm, dt, cs, theta_off, g, epi);
}

// Value of the log-density at theta.  Must follow dm_grad at the same theta (uses the c_uk it left in the
// scratch).  Terms are evaluated in T, accumulated in fp64; the two O(n ln n)-sized row terms are taken in
// fp64.  Every thread of the group receives the result.  Rec = CPsi<T> (element path) or PairRec<T> (grouped path):
// both keep the concentration of pair u*K + k in their member c.
template <typename T, int W, class Rec>
__device__ __noinline__ double dm_value(const Dims& dm, const DataTile<T>& dt, const ChainScratch& cs,
                                        uint32_t theta_off, const Group<W>& g) {
    const int tid = g.tid;
    constexpr int TS = Group<W>::SIZE;
    const int K = dm.K, D = dm.D, DK1 = dm.DK1, U = dt.U;
    const T* sigma = sp<T>(theta_off);
    const T* boff = sigma + D;
    const T* iraw = boff + DK1;
    const T* alpha = iraw + DK1;
    const Rec* cp = sp<Rec>(cs.cp);
    const int* rstart = sp<int>(dt.rstart);
    const int* grp = sp<int>(dt.grp);
    const T* ntot = sp<T>(dt.ntot);
    double lp = 0.0;
    // priors: HalfCauchy(0,1) on sigma_d (:703-707), Normal(0,1) on b_offset / ind_raw (:709-722), Normal(0,5) on
    // alpha (:736-741); the theta-independent constants live in dt.lconst
    {
        T acc = T(0);
#pragma unroll 1
        for (int i = D + tid; i < dm.P; i += TS) {
            const T v = sigma[i];                                   // theta[i]
            acc = fma(v * v, i < D + 2 * DK1 ? T(-0.5) : T(-0.02), acc);
        }
        lp += (double)acc;
    }
#pragma unroll 1
    for (int d = tid; d < D; d += TS) {
        // in the arithmetic type: the fp64 log1p of the fp32 build was 16 % of the NUTS kernel's instructions at
        // BASELINE config 3, executed by one warp while the others wait at the next barrier
        const T s = sigma[d];
        if (s >= T(0)) lp -= (double)t_log1p(s * s);
    }
    // DirichletMultinomial [TFP]: lbeta(c+y) - lbeta(c)                               (:746-751)
    // Guard (documented deviation): beyond c_max the lgamma differences have lost all
    // precision (the reference's fp64 graph then evaluates to the bare multinomial constant, a spurious mode
    // thousands of nats above the posterior that traps the chain); such states get a NaN value => rejected.
    // The limit is sccoda_sampler::reject_conc_above (default 1e6 in fp32, 1e15 in fp64; infinity switches the guard off).
    const T c_max = (T)dm.conc_max;
    // fp32: concentrations >= c_joint switch to the joint (cancellation-free) form, which already contains
    // -lgamma(c); never taken in the fp64 build
    const T c_joint = sizeof(T) == 4 ? T(SCCODA_C_JOINT) : T(SCCODA_C_JOINT_F64);
    bool need_fix = false;   // some concentration >= c_joint; made group-uniform below
    {
        T acc = T(0);
        bool bad = false;
#pragma unroll 1
        for (int e = tid; e < U * K; e += TS) {
            const int u = fast_div(e, dm.magicK);
            const T c = cp[e].c;
            if (c < c_joint) acc = fma(-(T)(rstart[u + 1] - rstart[u]), lgamma_pos(c), acc);
            else need_fix = true;
            bad = bad || !(c <= c_max);
        }
        lp += (double)acc;
        if (bad) lp = CUDART_NAN;
    }
#pragma unroll 1
    for (int n = tid; n < dm.N; n += TS) {
        if constexpr (Rec::has_totals) {
            // grouped path: the gradient pass left S_u in the record of the total pair
            if constexpr (sizeof(T) == 4) {
                // -[lgamma(S+n) - lgamma(S) - lgamma(n)] in the cancellation-free fp32 forms (as cc_value); the
                // sum of lgamma(n) was moved into the constant at staging
                const T S = cp[dt.tot0 + grp[n]].c, nn = ntot[n], rc = sp<T>(dt.nrc)[n];
                lp -= (double)(S >= T(6) ? lgamma_joint(S, nn, rc) : lgamma_shift(S, nn, rc) - lgamma_pos(S));
            } else {
                lp += lgamma_rowdiff_f64((double)cp[dt.tot0 + grp[n]].c, (double)ntot[n]);
            }
        } else {
            const Rec* cu = cp + grp[n] * K;
            // S in fp32 pairs, combined in fp64 (the conversions are the expensive part of an fp64 accumulation)
            T S0 = T(0), S1 = T(0), S2 = T(0), S3 = T(0);
            int k = 0;
    #pragma unroll 1
            for (; k + 3 < K; k += 4) {
                S0 += cu[k].c;
                S1 += cu[k + 1].c;
                S2 += cu[k + 2].c;
                S3 += cu[k + 3].c;
            }
    #pragma unroll 1
            for (; k < K; ++k) S0 += cu[k].c;
            if constexpr (sizeof(T) == 4) {
                // fp32 build: the cancellation-free joint form, as on the grouped layout (two fp64 logarithms per
                // row otherwise, on the critical path of the value pass)
                const T S = (S0 + S1) + (S2 + S3), nn = ntot[n], rc = sp<T>(dt.nrc)[n];
                lp -= (double)(S >= T(6) ? lgamma_joint(S, nn, rc) : lgamma_shift(S, nn, rc) - lgamma_pos(S));
            } else {
                lp += lgamma_rowdiff_f64(((double)S0 + (double)S1) + ((double)S2 + (double)S3), (double)ntot[n]);
            }
        }
    }
    // elements: lgamma(c + y) - lgamma(y); per-thread partial sums in T (<= a few dozen terms), total in fp64.
    // When some concentration is >= c_joint (uniform flag `fix`), the elements of those pairs take the joint form
    // lgamma(c+y) - lgamma(c) - lgamma(y) instead: summing the plain form in fp32 and correcting it afterwards would
    // leave the rounding of partial sums of size c ln c behind.
    const bool fix = sizeof(T) == 4 && g.any(need_fix);
    auto term_big = [&](T c, T y, T ycst) {
        return (fix && c >= c_joint) ? lgamma_joint(c, y, ycst) : lgamma_shift_big(c, y, ycst);
    };
    auto term_any = [&](T c, T y, T ycst) {
        return (fix && c >= c_joint) ? lgamma_joint(c, y, ycst) : lgamma_shift(c, y, ycst);
    };
    if constexpr (Rec::has_totals) {
        // grouped layout: counts >= 6 from the items already staged for the gradient (total pairs belong to the row
        // terms above; the Stirling remainder of a count is evaluated on the fly), counts < 6 from the tail
        // [nbig, N*K) of the element list
        T acc = T(0);
        const T* iy = sp<T>(dt.iy);
        const uint16_t* ipair = sp<uint16_t>(dt.ipair);
        const uint8_t* inreal = sp<uint8_t>(dt.inreal);
        const int NIs = dt.NImax;
#pragma unroll 1
        for (int i = tid; i < dt.NI; i += TS) {
            const int p = ipair[i];
            if (p >= dt.tot0) continue;
            const T c = cp[p].c;
            const int nr = inreal[i];
            if (fix && c >= c_joint) {
#pragma unroll 1
                for (int j = 0; j < nr; ++j) {
                    const T y = iy[j * NIs + i];
                    acc += lgamma_joint(c, y, stirling_rem_of(y));
                }
            } else {
#pragma unroll
                for (int j = 0; j < 5; ++j)
                    if (j < nr) acc += lgamma_shift_big_auto(c, iy[j * NIs + i]);
            }
        }
        const int ns = dm.NK - dt.nbig;
        if (dt.el_smem) {
            const Elem<T>* el = sp<Elem<T>>(dt.el);
            const T* eyc = sp<T>(dt.eyc);
#pragma unroll 1
            for (int e = tid; e < ns; e += TS) {
                const Elem<T> v = el[e];
                acc += term_any(cp[v.idx & 0xffffu].c, v.y, eyc[e]);
            }
        } else {
#pragma unroll 1
            for (int e = dt.nbig + tid; e < dm.NK; e += TS)
                acc += term_any(cp[__ldg(dt.eidx + e) & 0xffffu].c, __ldg(dt.ey + e), sizeof(T) == 4 ? __ldg(dt.eycst + e) : T(0));
        }
        lp += (double)acc;
    } else {
        T acc = T(0);
        if (dt.el_smem) {
            const Elem<T>* el = sp<Elem<T>>(dt.el);
            const T* eyc = sp<T>(dt.eyc);
#pragma unroll 2
            for (int e = tid; e < dt.nbig; e += TS) {
                const Elem<T> v = el[e];
                acc += term_big(cp[v.idx & 0xffffu].c, v.y, eyc[e]);
            }
#pragma unroll 1
            for (int e = dt.nbig + tid; e < dm.NK; e += TS) {
                const Elem<T> v = el[e];
                acc += term_any(cp[v.idx & 0xffffu].c, v.y, eyc[e]);   // mixed columns: branch on y
            }
        } else {
            // element list streamed from global memory (tiles too large for shared memory, fp64 parity build): same
            // split into counts >= 6 and mixed columns, independent loads in flight
            const T* __restrict__ ey = dt.ey;
            const T* __restrict__ eyc = dt.eycst;
            const uint32_t* __restrict__ eidx = dt.eidx;
#pragma unroll 8
            for (int e = tid; e < dt.nbig; e += TS)
                acc += term_big(cp[__ldg(eidx + e) & 0xffffu].c, __ldg(ey + e), sizeof(T) == 4 ? __ldg(eyc + e) : T(0));
#pragma unroll 1
            for (int e = dt.nbig + tid; e < dm.NK; e += TS)
                acc += term_any(cp[__ldg(eidx + e) & 0xffffu].c, __ldg(ey + e), sizeof(T) == 4 ? __ldg(eyc + e) : T(0));
        }
        lp += (double)acc;
    }
    lp = g.sum(lp) + dt.lconst;
    // HalfCauchy support: -inf as soon as one sigma_d < 0 [TFP]; every thread scans the D values (uniform)
    for (int d = 0; d < D; ++d)
        if (sigma[d] < T(0)) lp = -CUDART_INF;
    return lp;
}

}  // namespace sccoda
