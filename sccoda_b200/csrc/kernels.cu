// K2 / K3 — persistent HMC and NUTS kernels (one chain group per Markov chain, whole run in one launch)
// and the K1 unit-test kernel.  See include/sccoda_b200.h for the C ABI and DESIGN.md for the design.
//
// What the reference does with tfp.mcmc.sample_chain + HamiltonianMonteCarlo / NoUTurnSampler +
// Simple/DualAveragingStepSizeAdaptation (sccoda/model/scCODA_model.py:152-161, 276-315, 382-424, 495-553)
// happens here inside ONE kernel: momentum draw (Philox), leapfrog, Metropolis / multinomial tree
// sampling, step-size adaptation and trace write-out.  The chain state never leaves shared memory /
// registers; HBM sees only the draw write-out.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>

#include "../../include/sccoda_b200.h"
#include "kernels.h"
#include "model.cuh"
#include "philox.cuh"

namespace sccoda {

// ------------------------------------------------------------------------------------------------
// shared-memory plan (identical on host and device)
// ------------------------------------------------------------------------------------------------
__host__ __device__ inline size_t rnd16(size_t bytes) { return (bytes + 15) & ~size_t(15); }

struct SmemPlan {
    size_t y, ycst, ntot, xu, grp, rstart, data_total;                       // per CTA
    size_t c, psic, rowterm, G, Gu, beta, ind, gk, vec, red, chain_total;   // per chain (relative)
};

__host__ __device__ inline SmemPlan plan_smem(int N, int K, int D, int Umax, int nvec, int W, size_t ts) {
    SmemPlan p;
    const size_t NK = (size_t)N * K, UK = (size_t)Umax * K, P = (size_t)D + 2 * (size_t)D * (K - 1) + K;
    size_t o = 0;
    p.y = o; o += rnd16(NK * ts);
    p.ycst = o; o += (ts == 4) ? rnd16(NK * ts) : 0;  // fp64 build: the constant is identically 0, not staged
    p.ntot = o; o += rnd16((size_t)N * ts);
    p.xu = o; o += rnd16((size_t)Umax * D * ts);
    p.grp = o; o += rnd16((size_t)N * 4);
    p.rstart = o; o += rnd16((size_t)(Umax + 1) * 4);
    p.data_total = o;
    o = 0;
    p.red = o; o += rnd16((size_t)2 * W * 8);
    p.c = o; o += rnd16(UK * ts);
    p.psic = o; o += rnd16(UK * ts);
    p.rowterm = o; o += rnd16((size_t)N * ts);
    p.G = o; o += rnd16(NK * ts);
    p.Gu = o; o += rnd16(UK * ts);
    p.beta = o; o += rnd16((size_t)D * K * ts);
    p.ind = o; o += rnd16((size_t)D * (K - 1) * ts + ts);
    p.gk = o; o += rnd16((size_t)(D + 1) * K * ts);
    p.vec = o; o += rnd16((size_t)nvec * rnd16(P * ts));
    p.chain_total = o;
    return p;
}

template <typename T>
__device__ __forceinline__ size_t vec_stride(int P) {
    return rnd16((size_t)P * sizeof(T)) / sizeof(T);
}

// Stage one dataset into shared memory (whole CTA, coalesced; 128-bit when the tile is 16B aligned).
template <typename T>
__device__ __forceinline__ void copy_to_smem(T* dst, const T* __restrict__ src, int n) {
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && ((n * sizeof(T)) % 16 == 0);
    if (vec_ok) {
        const int nv = (int)(n * sizeof(T) / 16);
        const int4* s4 = reinterpret_cast<const int4*>(src);
        int4* d4 = reinterpret_cast<int4*>(dst);
        for (int i = threadIdx.x; i < nv; i += blockDim.x) d4[i] = __ldg(s4 + i);
    } else {
        for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
    }
}

template <typename T>
__device__ __forceinline__ void stage_dataset(const KArgs& a, int b, unsigned char* smem, const SmemPlan& pl,
                                              DataTile<T>& dt) {
    const int N = a.dm.N, K = a.dm.K, D = a.dm.D, Umax = a.Umax;
    T* y = reinterpret_cast<T*>(smem + pl.y);
    T* ycst = reinterpret_cast<T*>(smem + pl.ycst);
    T* ntot = reinterpret_cast<T*>(smem + pl.ntot);
    T* xu = reinterpret_cast<T*>(smem + pl.xu);
    int* grp = reinterpret_cast<int*>(smem + pl.grp);
    int* rstart = reinterpret_cast<int*>(smem + pl.rstart);
    copy_to_smem(y, static_cast<const T*>(a.y) + (size_t)b * N * K, N * K);
    if (sizeof(T) == 4) copy_to_smem(ycst, static_cast<const T*>(a.ycst) + (size_t)b * N * K, N * K);
    copy_to_smem(ntot, static_cast<const T*>(a.ntot) + (size_t)b * N, N);
    copy_to_smem(xu, static_cast<const T*>(a.xu) + (size_t)b * Umax * D, Umax * D);
    copy_to_smem(grp, a.grp + (size_t)b * N, N);
    copy_to_smem(rstart, a.rstart + (size_t)b * (Umax + 1), Umax + 1);
    dt.y = y; dt.ycst = ycst; dt.ntot = ntot; dt.xu = xu; dt.grp = grp; dt.rstart = rstart;
    dt.U = a.U[b];
    dt.ref = a.ref[b];
    dt.lconst = a.lconst[b];
}

template <typename T>
__device__ __forceinline__ void carve_chain(unsigned char* base, const SmemPlan& pl, ChainScratch<T>& cs, T*& vec,
                                            double*& red) {
    red = reinterpret_cast<double*>(base + pl.red);
    cs.c = reinterpret_cast<T*>(base + pl.c);
    cs.psic = reinterpret_cast<T*>(base + pl.psic);
    cs.rowterm = reinterpret_cast<T*>(base + pl.rowterm);
    cs.G = reinterpret_cast<T*>(base + pl.G);
    cs.Gu = reinterpret_cast<T*>(base + pl.Gu);
    cs.beta = reinterpret_cast<T*>(base + pl.beta);
    cs.ind = reinterpret_cast<T*>(base + pl.ind);
    cs.gk = reinterpret_cast<T*>(base + pl.gk);
    vec = reinterpret_cast<T*>(base + pl.vec);
}

// Common prologue: locate (dataset, chain), stage data, carve scratch.  Returns false for idle groups.
template <typename T, int W>
__device__ __forceinline__ bool prologue(const KArgs& a, int nvec, unsigned char* smem, DataTile<T>& dt,
                                         ChainScratch<T>& cs, T*& vec, Group<W>& g, int& b, int& c) {
    constexpr int TS = 32 * W;
    const int cpc = blockDim.x / TS;
    const int ctas_per_ds = (a.C + cpc - 1) / cpc;
    b = blockIdx.x / ctas_per_ds;
    const int cblk = blockIdx.x - b * ctas_per_ds;
    g.gid = threadIdx.x / TS;
    g.tid = threadIdx.x - g.gid * TS;
    c = cblk * cpc + g.gid;
    const SmemPlan pl = plan_smem(a.dm.N, a.dm.K, a.dm.D, a.Umax, nvec, W, sizeof(T));
    stage_dataset<T>(a, b, smem, pl, dt);
    carve_chain<T>(smem + pl.data_total + (size_t)g.gid * pl.chain_total, pl, cs, vec, g.red);
    // the reference column of beta is identically zero (scCODA_model.py:732-734); written once
    if (c < a.C)
        for (int d = g.tid; d < a.dm.D; d += TS) cs.beta[d * a.dm.K + dt.ref] = T(0);
    __syncthreads();
    return c < a.C;
}

// ------------------------------------------------------------------------------------------------
// K1 unit kernel: value + gradient at given states
// ------------------------------------------------------------------------------------------------
template <typename T, int W>
__global__ void __launch_bounds__(512) logp_grad_kernel(const __grid_constant__ KArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    DataTile<T> dt;
    ChainScratch<T> cs;
    T* vec;
    Group<W> g;
    int b, c;
    if (!prologue<T, W>(a, 2, smem, dt, cs, vec, g, b, c)) return;
    constexpr int TS = 32 * W;
    const int P = a.dm.P;
    const size_t vs = vec_stride<T>(P);
    T* th = vec;
    T* gr = vec + vs;
    const size_t chain = (size_t)b * a.C + c;
    const T* theta_in = static_cast<const T*>(a.init) + chain * P;
    for (int i = g.tid; i < P; i += TS) th[i] = theta_in[i];
    g.sync();
    const double lp = dm_eval<T, true, W>(a.dm, dt, cs, th, gr, g);
    T* grad_out = static_cast<T*>(a.draws) + chain * P;
    for (int i = g.tid; i < P; i += TS) grad_out[i] = gr[i];
    if (g.tid == 0) a.carry[chain] = lp;
}

// ------------------------------------------------------------------------------------------------
// step-size adaptation state (per chain, replicated in every thread of the group — all inputs uniform)
// ------------------------------------------------------------------------------------------------
template <typename T>
struct Adapt {
    T eps, H, log_avg, mu;
    int t;

    __device__ __forceinline__ void init(T eps0) {
        eps = eps0; H = T(0); log_avg = T(0); t = 0;
        mu = (T)log(10.0 * (double)eps0);
    }
    // [TFP] SimpleStepSizeAdaptation(adaptation_rate=0.01)
    __device__ __forceinline__ void simple(double lar, int step, int num_adapt, double log_target) {
        if (step < num_adapt) eps = (fmin(lar, 0.0) > log_target) ? eps * T(1.01) : eps / T(1.01);
    }
    // [TFP] DualAveragingStepSizeAdaptation(gamma=0.05, t0=10, kappa=0.75, mu=log(10 eps0))
    __device__ __forceinline__ void dual(double log_accept_prob, int num_adapt, T target) {
        if (t >= num_adapt) return;
        t += 1;
        const T a = (log_accept_prob > -CUDART_INF) ? (T)exp(log_accept_prob) : T(0);
        H += target - a;
        const T tt = (T)t;
        const T log_eps = mu - H * sqrt(tt) / ((tt + T(10)) * T(0.05));
        const T eta = (T)pow((double)tt, -0.75);
        log_avg = eta * log_eps + (T(1) - eta) * log_avg;
        eps = (t < num_adapt) ? (T)exp((double)log_eps) : (T)exp((double)log_avg);
    }
};

template <typename T>
__device__ __forceinline__ void write_stats(T* st, double lp, double lar, bool acc, T eps, bool div, int leap,
                                            double energy, bool maxd) {
    st[SCCODA_ST_TARGET_LOG_PROB] = (T)lp;
    st[SCCODA_ST_LOG_ACCEPT_RATIO] = (T)lar;
    st[SCCODA_ST_IS_ACCEPTED] = acc ? T(1) : T(0);
    st[SCCODA_ST_STEP_SIZE] = eps;
    st[SCCODA_ST_DIVERGING] = div ? T(1) : T(0);
    st[SCCODA_ST_LEAPFROGS_TAKEN] = (T)leap;
    st[SCCODA_ST_ENERGY] = (T)energy;
    st[SCCODA_ST_REACH_MAX_DEPTH] = maxd ? T(1) : T(0);
}

template <typename T>
__device__ __forceinline__ void load_chain_state(const KArgs& a, size_t chain, T* th, Adapt<T>& ad, int tid, int TS) {
    const int P = a.dm.P;
    if (a.step_begin == 0) {
        const T* init = static_cast<const T*>(a.init) + chain * P;
        for (int i = tid; i < P; i += TS) th[i] = init[i];
        ad.init((T)a.step_size);
    } else {
        const double* cr = a.carry + chain * (P + SCCODA_NCARRY);
        for (int i = tid; i < P; i += TS) th[i] = (T)cr[i];
        ad.init((T)a.step_size);
        ad.eps = (T)cr[P + 0]; ad.H = (T)cr[P + 1]; ad.log_avg = (T)cr[P + 2]; ad.t = (int)cr[P + 3];
    }
}
template <typename T>
__device__ __forceinline__ void store_chain_state(const KArgs& a, size_t chain, const T* th, const Adapt<T>& ad,
                                                  double lp, int tid, int TS) {
    if (a.carry == nullptr) return;
    const int P = a.dm.P;
    double* cr = a.carry + chain * (P + SCCODA_NCARRY);
    for (int i = tid; i < P; i += TS) cr[i] = (double)th[i];
    if (tid == 0) {
        cr[P + 0] = (double)ad.eps; cr[P + 1] = (double)ad.H; cr[P + 2] = (double)ad.log_avg; cr[P + 3] = (double)ad.t;
        cr[P + 4] = lp; cr[P + 5] = 0.0; cr[P + 6] = 0.0; cr[P + 7] = 0.0;
    }
}

// ------------------------------------------------------------------------------------------------
// K2: HMC  ([TFP] HamiltonianMonteCarlo.one_step + MetropolisHastings + step-size adaptation)
// ------------------------------------------------------------------------------------------------
template <typename T, int W>
__global__ void __launch_bounds__(512) hmc_kernel(const __grid_constant__ KArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    DataTile<T> dt;
    ChainScratch<T> cs;
    T* vec;
    Group<W> g;
    int b, c;
    if (!prologue<T, W>(a, 5, smem, dt, cs, vec, g, b, c)) return;
    constexpr int TS = 32 * W;
    const int P = a.dm.P, tid = g.tid;
    const size_t vs = vec_stride<T>(P);
    T* thc = vec;            // current state
    T* gc = vec + vs;        // its gradient
    T* thw = vec + 2 * vs;   // leapfrog position
    T* gw = vec + 3 * vs;    // its gradient
    T* mom = vec + 4 * vs;   // momentum
    const size_t chain = (size_t)b * a.C + c;
    const RngKey key{(uint32_t)(a.seed & 0xffffffffu), (uint32_t)(a.seed >> 32), a.chain_id0 + (uint32_t)chain};

    Adapt<T> ad;
    load_chain_state<T>(a, chain, thc, ad, tid, TS);
    g.sync();
    double lp_cur = dm_eval<T, true, W>(a.dm, dt, cs, thc, gc, g);

    const int L = a.num_leapfrog;
    const double log_target = log(a.target_accept);
    T* draws = static_cast<T*>(a.draws) + chain * (size_t)a.num_results * P;
    T* stats = static_cast<T*>(a.stats) + chain * (size_t)a.num_results * SCCODA_NSTAT;

    for (int t = a.step_begin; t < a.step_end; ++t) {
        const T eps = ad.eps;
        T kin = T(0);
        for (int i = tid; i < P; i += TS) {
            const T p = rng_normal<T>(key, (uint32_t)i, (uint32_t)t);
            kin += p * p;
            thw[i] = thc[i];
            mom[i] = fma(T(0.5) * eps, gc[i], p);
        }
        double lp_new = 0.0;
        for (int l = 0; l < L; ++l) {
            for (int i = tid; i < P; i += TS) thw[i] = fma(eps, mom[i], thw[i]);
            g.sync();
            if (l == L - 1) lp_new = dm_eval<T, true, W>(a.dm, dt, cs, thw, gw, g);
            else dm_eval<T, false, W>(a.dm, dt, cs, thw, gw, g);
            const T f = (l == L - 1) ? T(0.5) * eps : eps;
            for (int i = tid; i < P; i += TS) mom[i] = fma(f, gw[i], mom[i]);
        }
        for (int i = tid; i < P; i += TS) kin -= mom[i] * mom[i];
        const double dkin = g.sum((double)kin);
        // [TFP] log_accept_ratio = safe_sum([logp', -logp, kinetic - kinetic']); non-finite -> -inf
        double lar = lp_new - lp_cur + 0.5 * dkin;
        if (!isfinite(lar)) lar = -CUDART_INF;
        const double u = (double)rng_uniform<T>(key, 0u, (uint32_t)t, STREAM_ACCEPT);
        const bool acc = log(u) < lar;
        if (acc) {
            T* tmp = thc; thc = thw; thw = tmp;
            tmp = gc; gc = gw; gw = tmp;
            lp_cur = lp_new;
        }
        T traced_eps = eps;
        if (a.method == SCCODA_METHOD_HMC) {
            ad.simple(lar, t, a.num_adapt, log_target);
        } else {
            ad.dual(fmin(lar, 0.0), a.num_adapt, (T)a.target_accept);
            traced_eps = (T)exp((double)ad.log_avg);  // scCODA_model.py:408,419
        }
        const int r = t - a.num_burnin;
        if (r >= 0 && a.store_draws) {
            T* dr = draws + (size_t)r * P;
            for (int i = tid; i < P; i += TS) dr[i] = thc[i];
            if (tid == 0)
                write_stats<T>(stats + (size_t)r * SCCODA_NSTAT, lp_cur, lar, acc, traced_eps, lar < -1000.0, L,
                               CUDART_NAN, false);
        }
    }
    store_chain_state<T>(a, chain, thc, ad, lp_cur, tid, TS);
}

// ------------------------------------------------------------------------------------------------
// K3: NUTS  ([TFP] NoUTurnSampler.one_step: iterative doubling, multinomial sampling, generalized U-turn,
//            checkpoint memory; restated in SURVEY §8c) + dual averaging
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double logaddexp_d(double x, double y) {
    if (x == -CUDART_INF) return y;
    if (y == -CUDART_INF) return x;
    const double m = fmax(x, y);
    return m + log(exp(x - m) + exp(y - m));
}

template <typename T, int W>
__global__ void __launch_bounds__(512) nuts_kernel(const __grid_constant__ KArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    DataTile<T> dt;
    ChainScratch<T> cs;
    T* vec;
    Group<W> g;
    int b, c;
    const int maxd = a.max_depth;
    if (!prologue<T, W>(a, 15 + 2 * (maxd + 1), smem, dt, cs, vec, g, b, c)) return;
    constexpr int TS = 32 * W;
    const int P = a.dm.P, tid = g.tid;
    const size_t vs = vec_stride<T>(P);
    T* cq = vec + 0 * vs;   // candidate (= current state between transitions)
    T* cg = vec + 1 * vs;
    T* lq = vec + 2 * vs;   // left end  (q, p, grad)
    T* lpm = vec + 3 * vs;
    T* lg = vec + 4 * vs;
    T* rq = vec + 5 * vs;   // right end
    T* rpm = vec + 6 * vs;
    T* rg = vec + 7 * vs;
    T* sq = vec + 8 * vs;   // subtree candidate
    T* sg = vec + 9 * vs;
    T* rho = vec + 10 * vs;
    T* rho_s = vec + 11 * vs;
    T* q = vec + 12 * vs;   // leaf being integrated
    T* p = vec + 13 * vs;
    T* gq = vec + 14 * vs;
    T* ck_p = vec + 15 * vs;
    T* ck_rho = vec + (15 + (maxd + 1)) * vs;
    const size_t chain = (size_t)b * a.C + c;
    const RngKey key{(uint32_t)(a.seed & 0xffffffffu), (uint32_t)(a.seed >> 32), a.chain_id0 + (uint32_t)chain};

    Adapt<T> ad;
    load_chain_state<T>(a, chain, cq, ad, tid, TS);
    g.sync();
    double lp_cur = dm_eval<T, true, W>(a.dm, dt, cs, cq, cg, g);

    T* draws = static_cast<T*>(a.draws) + chain * (size_t)a.num_results * P;
    T* stats = static_cast<T*>(a.stats) + chain * (size_t)a.num_results * SCCODA_NSTAT;

    for (int t = a.step_begin; t < a.step_end; ++t) {
        const T eps = ad.eps;
        T kin = T(0);
        for (int i = tid; i < P; i += TS) {
            const T pm = rng_normal<T>(key, (uint32_t)i, (uint32_t)t);
            kin += pm * pm;
            lq[i] = cq[i]; lpm[i] = pm; lg[i] = cg[i];
            rq[i] = cq[i]; rpm[i] = pm; rg[i] = cg[i];
            rho[i] = pm;
        }
        const double e0 = lp_cur - 0.5 * g.sum((double)kin);
        double l_lp = lp_cur, r_lp = lp_cur, c_lp = lp_cur, c_energy = e0, s_lp = lp_cur, s_en = e0;
        double wsum = 0.0, accept_sum = 0.0;
        int leapfrogs = 0, depth = 0;
        bool is_acc = false, cont = true, not_div = true;
        uint32_t leaf_serial = 0;
        while (depth < maxd && cont) {
            const bool forward = rng_bit(key, (uint32_t)depth, (uint32_t)t, STREAM_NUTS_DIR) != 0u;
            T* eq = forward ? rq : lq;
            T* ep = forward ? rpm : lpm;
            T* eg = forward ? rg : lg;
            const T se = forward ? eps : -eps;
            double lqv = forward ? r_lp : l_lp;
            for (int i = tid; i < P; i += TS) {
                q[i] = eq[i]; p[i] = ep[i]; gq[i] = eg[i];
                rho_s[i] = T(0);
            }
            const unsigned n_leaves = 1u << depth;
            double w_s = -CUDART_INF;
            bool alive = true;
            for (unsigned leaf = 0; leaf < n_leaves && alive; ++leaf) {
                for (int i = tid; i < P; i += TS) {
                    const T ph = fma(T(0.5) * se, gq[i], p[i]);
                    p[i] = ph;
                    q[i] = fma(se, ph, q[i]);
                }
                g.sync();
                lqv = dm_eval<T, true, W>(a.dm, dt, cs, q, gq, g);
                T kk = T(0);
                for (int i = tid; i < P; i += TS) {
                    const T pn = fma(T(0.5) * se, gq[i], p[i]);
                    p[i] = pn;
                    rho_s[i] += pn;
                    kk += pn * pn;
                }
                leapfrogs += 1;
                const int idx_max = __popc(leaf >> 1);
                bool no_uturn = true;
                if ((leaf & 1u) == 0u) {
                    T* cp = ck_p + (size_t)idx_max * vs;
                    T* cr = ck_rho + (size_t)idx_max * vs;
                    for (int i = tid; i < P; i += TS) { cp[i] = p[i]; cr[i] = rho_s[i]; }
                } else {
                    const int n_chk = __ffs(~leaf) - 1;  // trailing ones
                    for (int s = idx_max; s > idx_max - n_chk && no_uturn; --s) {
                        const T* cp = ck_p + (size_t)s * vs;
                        const T* cr = ck_rho + (size_t)s * vs;
                        T d1 = T(0), d2 = T(0);
                        for (int i = tid; i < P; i += TS) {
                            const T sp = rho_s[i] - cr[i] + cp[i];
                            d1 = fma(sp, cp[i], d1);
                            d2 = fma(sp, p[i], d2);
                        }
                        const double D1 = g.sum((double)d1), D2 = g.sum((double)d2);
                        if (!(D1 >= 0.0 && D2 >= 0.0)) no_uturn = false;
                    }
                }
                double energy = lqv - 0.5 * g.sum((double)kk);
                if (isnan(energy)) energy = -CUDART_INF;
                const double delta = energy - e0;
                const bool divergent = !(-delta < 1000.0);
                const double w_new = logaddexp_d(w_s, delta);
                const double u = (double)rng_uniform<T>(key, leaf_serial++, (uint32_t)t, STREAM_NUTS_LEAF);
                const double thresh = (w_new > -CUDART_INF) ? delta - w_new : CUDART_NAN;
                if (log1p(-u) <= thresh) {
                    for (int i = tid; i < P; i += TS) { sq[i] = q[i]; sg[i] = gq[i]; }
                    s_lp = lqv; s_en = energy;
                }
                w_s = w_new;
                accept_sum += exp(fmin(delta, 0.0));
                if (divergent) not_div = false;
                alive = (!divergent) && no_uturn;
            }
            const double tree_w = alive ? w_s : -CUDART_INF;
            double thresh = tree_w - wsum;
            if (isnan(thresh)) thresh = 0.0;
            const double u = (double)rng_uniform<T>(key, (uint32_t)depth, (uint32_t)t, STREAM_NUTS_TOP);
            if ((log1p(-u) <= thresh) && alive) {
                for (int i = tid; i < P; i += TS) { cq[i] = sq[i]; cg[i] = sg[i]; }
                c_lp = s_lp; c_energy = s_en; is_acc = true;
            }
            wsum = logaddexp_d(wsum, tree_w);
            T d1 = T(0), d2 = T(0);
            for (int i = tid; i < P; i += TS) {
                eq[i] = q[i]; ep[i] = p[i]; eg[i] = gq[i];
                const T r = rho[i] + rho_s[i];
                rho[i] = r;
                d1 = fma(r, lpm[i], d1);   // ep aliases lpm or rpm: already updated for this thread's i
                d2 = fma(r, rpm[i], d2);
            }
            if (forward) r_lp = lqv; else l_lp = lqv;
            const double D1 = g.sum((double)d1), D2 = g.sum((double)d2);
            cont = alive && (D1 >= 0.0) && (D2 >= 0.0);
            depth += 1;
        }
        lp_cur = c_lp;
        const double lar = accept_sum > 0.0 ? log(accept_sum / (double)leapfrogs) : -CUDART_INF;
        ad.dual(fmin(lar, 0.0), a.num_adapt, (T)a.target_accept);
        const int r = t - a.num_burnin;
        if (r >= 0 && a.store_draws) {
            T* dr = draws + (size_t)r * P;
            for (int i = tid; i < P; i += TS) dr[i] = cq[i];
            if (tid == 0)
                write_stats<T>(stats + (size_t)r * SCCODA_NSTAT, lp_cur, lar, is_acc, eps, !not_div, leapfrogs,
                               c_energy, cont);
        }
    }
    store_chain_state<T>(a, chain, cq, ad, lp_cur, tid, TS);
}

// ------------------------------------------------------------------------------------------------
// launch helpers
// ------------------------------------------------------------------------------------------------
template <typename T, int W>
static cudaError_t launch_one(int kind, const KArgs& a, int grid, int block, size_t smem, cudaStream_t st) {
    cudaError_t e;
    if (kind == KIND_LOGP) {
        e = cudaFuncSetAttribute(logp_grad_kernel<T, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        logp_grad_kernel<T, W><<<grid, block, smem, st>>>(a);
    } else if (kind == KIND_HMC) {
        e = cudaFuncSetAttribute(hmc_kernel<T, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        hmc_kernel<T, W><<<grid, block, smem, st>>>(a);
    } else {
        e = cudaFuncSetAttribute(nuts_kernel<T, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        nuts_kernel<T, W><<<grid, block, smem, st>>>(a);
    }
    return cudaGetLastError();
}

template <typename T>
static cudaError_t launch_w(int kind, int W, const KArgs& a, int grid, int block, size_t smem, cudaStream_t st) {
    switch (W) {
        case 1: return launch_one<T, 1>(kind, a, grid, block, smem, st);
        case 2: return launch_one<T, 2>(kind, a, grid, block, smem, st);
        case 4: return launch_one<T, 4>(kind, a, grid, block, smem, st);
        case 8: return launch_one<T, 8>(kind, a, grid, block, smem, st);
        case 16: return launch_one<T, 16>(kind, a, grid, block, smem, st);
        default: return cudaErrorInvalidValue;
    }
}

int nvec_for(int kind, int max_depth) {
    if (kind == KIND_LOGP) return 2;
    if (kind == KIND_HMC) return 5;
    return 15 + 2 * (max_depth + 1);
}

size_t smem_bytes_for(const KArgs& a, int kind, int W, int chains_per_cta, int dtype) {
    const size_t ts = dtype == SCCODA_F64 ? 8 : 4;
    const SmemPlan pl = plan_smem(a.dm.N, a.dm.K, a.dm.D, a.Umax, nvec_for(kind, a.max_depth), W, ts);
    return pl.data_total + (size_t)chains_per_cta * pl.chain_total;
}

cudaError_t launch_kernel(int kind, int dtype, int W, int chains_per_cta, const KArgs& a, cudaStream_t st) {
    const int ctas_per_ds = (a.C + chains_per_cta - 1) / chains_per_cta;
    const int grid = a.B * ctas_per_ds;
    const int block = chains_per_cta * 32 * W;
    const size_t smem = smem_bytes_for(a, kind, W, chains_per_cta, dtype);
    if (dtype == SCCODA_F64) return launch_w<double>(kind, W, a, grid, block, smem, st);
    return launch_w<float>(kind, W, a, grid, block, smem, st);
}

}  // namespace sccoda
