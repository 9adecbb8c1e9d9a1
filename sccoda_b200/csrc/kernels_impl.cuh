// K2 / K3 — persistent HMC and NUTS kernels (one chain group per Markov chain, whole run in one launch)
// and the K1 unit-test kernel.  Included by kernels_f32.cu / kernels_f64.cu, which instantiate one dtype each.
//
// What the reference does with tfp.mcmc.sample_chain + HamiltonianMonteCarlo / NoUTurnSampler +
// Simple/DualAveragingStepSizeAdaptation (sccoda/model/scCODA_model.py:152-161, 276-315, 382-424, 495-553)
// happens here inside ONE kernel: momentum draw (Philox), leapfrog, Metropolis / multinomial tree
// sampling, step-size adaptation and trace write-out.  The chain state never leaves shared memory /
// registers; HBM sees only the draw write-out.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/sccoda_b200.h"
#include "kernels.h"
#include "grouped.cuh"
#include "model.cuh"
#include "philox.cuh"

namespace sccoda {

// ------------------------------------------------------------------------------------------------
// shared-memory carving (plan_smem is shared with the host, kernels.h)
// ------------------------------------------------------------------------------------------------
// byte stride between consecutive state vectors of one chain
template <typename T>
__device__ __forceinline__ uint32_t vec_stride_bytes(int P) {
    return (uint32_t)rnd16((size_t)P * sizeof(T));
}

template <typename V>
__device__ __forceinline__ void copy_to_smem(V* dst, const V* __restrict__ src, int n) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
}

template <typename T>
__device__ __forceinline__ void stage_dataset(const KArgs& a, int b, const SmemPlan& pl, DataTile<T>& dt) {
    const int N = a.dm.N, K = a.dm.K, D = a.dm.D, Umax = a.Umax, NK = a.dm.NK;
    dt.ntot = (uint32_t)pl.ntot; dt.xu = (uint32_t)pl.xu; dt.grp = (uint32_t)pl.grp; dt.rstart = (uint32_t)pl.rstart;
    dt.colperm = (uint32_t)pl.colperm; dt.el = (uint32_t)pl.el; dt.eyc = (uint32_t)pl.eyc;
    dt.el_smem = a.el_smem;
    dt.ey = static_cast<const T*>(a.ey) + (size_t)b * NK;
    dt.eidx = a.eidx + (size_t)b * NK;
    dt.eycst = static_cast<const T*>(a.eycst) + (size_t)b * NK;
    copy_to_smem(sp<T>(dt.ntot), static_cast<const T*>(a.ntot) + (size_t)b * N, N);
    copy_to_smem(sp<T>(dt.xu), static_cast<const T*>(a.xu) + (size_t)b * Umax * D, Umax * D);
    copy_to_smem(sp<int>(dt.grp), a.grp + (size_t)b * N, N);
    copy_to_smem(sp<int>(dt.rstart), a.rstart + (size_t)b * (Umax + 1), Umax + 1);
    copy_to_smem(sp<int>(dt.colperm), a.colperm + (size_t)b * K, K);
    dt.nbig = a.nbig[b];
    if (a.el_smem) {
        // (count, index) records + host constants; with the grouped layout only the tail of counts below 6 (the value
        // pass reads the counts >= 6 from the items)
        Elem<T>* el = sp<Elem<T>>(dt.el);
        T* eyc = sp<T>(dt.eyc);
        const int e0 = a.grouped ? dt.nbig : 0;
        for (int e = threadIdx.x; e < NK - e0; e += blockDim.x) {   // coalesced reads
            Elem<T> v;
            v.y = dt.ey[e0 + e];
            v.idx = dt.eidx[e0 + e];
            el[e] = v;
            eyc[e] = dt.eycst[e0 + e];
        }
    }
    dt.U = a.U[b];
    dt.ref = a.ref[b];
    dt.lconst = a.lconst[b];
    dt.pcoef = (uint32_t)pl.pcoef; dt.iy = (uint32_t)pl.iy; dt.ipair = (uint32_t)pl.ipair; dt.ipos = (uint32_t)pl.ipos;
    dt.inreal = (uint32_t)pl.inreal;
    dt.oy = nullptr; dt.opair = nullptr; dt.opos = nullptr; dt.odesc = nullptr;
    dt.NI = 0; dt.NG = 0; dt.NImax = a.NImax; dt.NF = a.NF; dt.tot0 = Umax * K;
    dt.nrc = (uint32_t)pl.nrc;
    if (sizeof(T) == 4) {
        // value pass, row terms in fp32: per-row constants r(n) / lgamma(n) in fp64 here, sum_n lgamma(n) into the
        // constant (one partial per warp, added in warp order: independent of everything but the block size)
        T* nrc = sp<T>(dt.nrc);
        double* red = sp<double>((uint32_t)pl.stage_red);
        const T* ntot = static_cast<const T*>(a.ntot) + (size_t)b * N;
        double lg_sum = 0.0;
        for (int n = threadIdx.x; n < N; n += blockDim.x) {
            const double nn = (double)ntot[n];
            const double lg = lgamma(nn);
            nrc[n] = (T)(nn >= 6.0 ? lg - ((nn - 0.5) * log(nn) - nn + 0.91893853320467274178) : lg);
            lg_sum += lg;
        }
        for (int o = 16; o > 0; o >>= 1) lg_sum += __shfl_xor_sync(0xffffffffu, lg_sum, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = lg_sum;
        __syncthreads();
        double tot = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[w];
        dt.lconst -= tot;
    }
    if (a.grouped) {
        const int NP = Umax * (K + 1);
        copy_to_smem(sp<T>(dt.pcoef), static_cast<const T*>(a.pcoef) + (size_t)b * NP * 6, NP * 6);
        copy_to_smem(sp<T>(dt.iy), static_cast<const T*>(a.iy) + (size_t)b * kItemR * a.NImax, kItemR * a.NImax);
        copy_to_smem(sp<uint16_t>(dt.ipair), a.ipair + (size_t)b * a.NImax, a.NImax);
        copy_to_smem(sp<uint16_t>(dt.ipos), a.ipos + (size_t)b * a.NImax, a.NImax);
        copy_to_smem(sp<uint8_t>(dt.inreal), a.inreal + (size_t)b * a.NImax, a.NImax);
        dt.oy = static_cast<const T*>(a.oy) + (size_t)b * a.NOmax;
        dt.opair = a.opair + (size_t)b * a.NGmax;
        dt.opos = a.opos + (size_t)b * a.NGmax;
        dt.odesc = a.odesc + (size_t)b * a.NGmax;
        dt.NI = a.NI[b];
        dt.NG = a.NG[b];
    }
}

__device__ __forceinline__ void carve_chain(uint32_t base, const SmemPlan& pl, ChainScratch& cs, uint32_t& red) {
    red = base + (uint32_t)pl.red;
    cs.cp = base + (uint32_t)pl.cp;
    cs.rowterm = base + (uint32_t)pl.rowterm;
    cs.G = base + (uint32_t)pl.G;
    cs.Gu = base + (uint32_t)pl.Gu;
    cs.beta = base + (uint32_t)pl.beta;
    cs.ind = base + (uint32_t)pl.ind;
    cs.gk = base + (uint32_t)pl.gk;
    cs.vec = base + (uint32_t)pl.vec;
}

// Common prologue: locate (dataset, chain), stage data, carve scratch.
// Returns false for idle groups (chain index past C).
template <typename T, int W>
__device__ __forceinline__ bool prologue(const KArgs& a, int nvec, DataTile<T>& dt, ChainScratch& cs, Group<W>& g,
                                         int& b, int& c) {
    constexpr int TS = 32 * W;
    const int cpc = blockDim.x / TS;
    const int ctas_per_ds = (a.C + cpc - 1) / cpc;
    b = blockIdx.x / ctas_per_ds;
    const int cblk = blockIdx.x - b * ctas_per_ds;
    g.gid = threadIdx.x / TS;
    g.tid = threadIdx.x - g.gid * TS;
    c = cblk * cpc + g.gid;
    const SmemPlan pl = plan_smem(a.dm.N, a.dm.K, a.dm.D, a.Umax, a.el_smem ? a.el_count : 0, a.grouped ? a.NImax : 0, a.NF, nvec, W,
                                  sizeof(T));
    stage_dataset<T>(a, b, pl, dt);
    carve_chain((uint32_t)(pl.data_total + (size_t)g.gid * pl.chain_total), pl, cs, g.red);
    // the reference column of beta is identically zero (scCODA_model.py:732-734); written once
    if (c < a.C) {
        for (int d = g.tid; d < a.dm.D; d += TS) sp<T>(cs.beta)[d * a.dm.K + dt.ref] = T(0);
        // grouped path: result slots that no item / odd group owns stay zero for the whole run
        if (a.grouped)
            for (int i = g.tid; i < a.Umax * (a.dm.K + 1) * a.NF + 1; i += TS) sp<T>(cs.G)[i] = T(0);
    }
    __syncthreads();
    return c < a.C;
}

#define SCCODA_BOUNDS(W) __launch_bounds__((W) <= 4 ? 128 : 32 * (W), (W) <= 4 ? 7 : 1)

// Gradient / value of the path the kernel was instantiated for (GR: grouped pair/item layout).
// hands a zero gradient to the epilogue for the frozen parameters of a SimpleModel run (Dims::freeze)
template <typename T, class Epi>
struct FrozenEpi {
    const Epi& epi;
    const Dims& dm;
    __device__ __forceinline__ void operator()(int i, T gr) const { epi(i, frozen_param(dm, i) ? T(0) : gr); }
};

template <typename T, int W, bool GR, class Epi>
__device__ __forceinline__ void grad_of(const Dims& dm, const DataTile<T>& dt, const ChainScratch& cs, uint32_t theta_off,
                                        const Group<W>& g, const Epi& epi) {
    const FrozenEpi<T, Epi> fe{epi, dm};
    if constexpr (GR) dm_grad_items<T, W>(dm, dt, cs, theta_off, g, fe);
    else dm_grad<T, W>(dm, dt, cs, theta_off, g, fe);
}
template <typename T, int W, bool GR>
__device__ __forceinline__ double value_of(const Dims& dm, const DataTile<T>& dt, const ChainScratch& cs,
                                           uint32_t theta_off, const Group<W>& g) {
    if constexpr (GR) return dm_value<T, W, PairRec<T>>(dm, dt, cs, theta_off, g);
    else return dm_value<T, W, CPsi<T>>(dm, dt, cs, theta_off, g);
}

// Gradient and value at the same state (a NUTS leaf).  With at least eight warps per chain the second half of the
// gradient pass (items / elements, brackets or column sums, chain rule) and the value pass run SIDE BY SIDE on two
// parts of the chain's warps: both only read what the first half wrote, and a chain with many warps is
// bound by the latency of its dependent instruction chains, not by issue slots (BASELINE config 3: 16 warps per
// chain on one SM).  Each half synchronises on its own named barrier and reduces in its own half of the scratch.
template <typename T, int W, bool GR, class Epi>
__device__ __forceinline__ double grad_value_of(const Dims& dm, const DataTile<T>& dt, const ChainScratch& cs,
                                                uint32_t theta_off, const Group<W>& g, const Epi& epi) {
    if constexpr (W >= 8) {
        // the value pass is the longer of the two: it gets 5/8 of the warps
        constexpr int HB = (3 * W) / 8, HV = W - HB;
        const FrozenEpi<T, Epi> fe{epi, dm};
        if constexpr (GR) dm_grad_items_front<T, W>(dm, dt, cs, theta_off, g);
        else dm_grad_front<T, W>(dm, dt, cs, theta_off, g);
        const bool second = g.tid >= 32 * HB;
        double* out = sp<double>(cs.Gu);                           // free once the front half is done
        if (!second) {
            Group<HB> h;
            h.tid = g.tid;
            h.gid = 8 + 2 * g.gid;                                 // named barriers 9, 10 (the chain groups use 1..4)
            h.red = g.red;                                         // the group's scratch is double[2 W]
            if constexpr (GR) dm_grad_items_back<T, HB>(dm, dt, cs, theta_off, h, fe);
            else dm_grad_back<T, HB>(dm, dt, cs, theta_off, h, fe);
        } else {
            Group<HV> h;
            h.tid = g.tid - 32 * HB;
            h.gid = 9 + 2 * g.gid;
            h.red = g.red + 16u * HB;
            double v;
            if constexpr (GR) v = dm_value<T, HV, PairRec<T>>(dm, dt, cs, theta_off, h);
            else v = dm_value<T, HV, CPsi<T>>(dm, dt, cs, theta_off, h);
            if (h.tid == 0) *out = v;
        }
        g.sync();
        const double v = *out;
        g.sync();                                                  // `out` is reused by the next gradient's row sums
        return v;
    } else {
        grad_of<T, W, GR>(dm, dt, cs, theta_off, g, epi);
        return value_of<T, W, GR>(dm, dt, cs, theta_off, g);
    }
}

// ------------------------------------------------------------------------------------------------
// K1 unit kernel: value + gradient at given states
// ------------------------------------------------------------------------------------------------
template <typename T, int W, bool GR>
__global__ void SCCODA_BOUNDS(W) logp_grad_kernel(const __grid_constant__ KArgs a) {
    DataTile<T> dt;
    ChainScratch cs;
    Group<W> g;
    int b, c;
    if (!prologue<T, W>(a, 2, dt, cs, g, b, c)) return;
    constexpr int TS = 32 * W;
    const int P = a.dm.P;
    const uint32_t th = cs.vec, gr = cs.vec + vec_stride_bytes<T>(P);
    const size_t chain = (size_t)b * a.C + c;
    const T* theta_in = static_cast<const T*>(a.init) + chain * P;
    for (int i = g.tid; i < P; i += TS) sp<T>(th)[i] = theta_in[i];
    g.sync();
    grad_of<T, W, GR>(a.dm, dt, cs, th, g, StoreGrad<T>{sp<T>(gr)});
    const double lp = value_of<T, W, GR>(a.dm, dt, cs, th, g);
    T* grad_out = static_cast<T*>(a.draws) + chain * P;
    for (int i = g.tid; i < P; i += TS) grad_out[i] = sp<T>(gr)[i];
    if (g.tid == 0) a.carry[chain] = lp;
}

// ------------------------------------------------------------------------------------------------
// step-size adaptation state (per chain, replicated in every thread of the group — all inputs uniform)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float t_exp(float v) { return expf(v); }
__device__ __forceinline__ double t_exp(double v) { return exp(v); }
__device__ __forceinline__ float t_log(float v) { return logf(v); }
__device__ __forceinline__ double t_log(double v) { return log(v); }
__device__ __forceinline__ float t_pow_m075(float t) { return expf(-0.75f * logf(t)); }
__device__ __forceinline__ double t_pow_m075(double t) { return pow(t, -0.75); }

template <typename T>
struct Adapt {
    T eps, H, log_avg, mu;
    int t;

    __device__ __forceinline__ void init(T eps0) {
        eps = eps0; H = T(0); log_avg = T(0); t = 0;
        mu = t_log(T(10) * eps0);
    }
    // [TFP] SimpleStepSizeAdaptation(adaptation_rate=0.01)
    __device__ __forceinline__ void simple(double lar, int step, int num_adapt, double log_target) {
        if (step < num_adapt) eps = (fmin(lar, 0.0) > log_target) ? eps * T(1.01) : eps / T(1.01);
    }
    // [TFP] DualAveragingStepSizeAdaptation(gamma=0.05, t0=10, kappa=0.75, mu=log(10 eps0))
    __device__ __forceinline__ void dual(double log_accept_prob, int num_adapt, T target) {
        if (t >= num_adapt) return;
        t += 1;
        const T a = (log_accept_prob > -CUDART_INF) ? t_exp((T)log_accept_prob) : T(0);
        H += target - a;
        const T tt = (T)t;
        const T log_eps = mu - H * sqrt(tt) / ((tt + T(10)) * T(0.05));
        const T eta = t_pow_m075(tt);
        log_avg = eta * log_eps + (T(1) - eta) * log_avg;
        eps = (t < num_adapt) ? t_exp(log_eps) : t_exp(log_avg);
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
    ad.init((T)a.step_size);
    if (a.step_begin == 0) {
        const T* init = static_cast<const T*>(a.init) + chain * P;
        for (int i = tid; i < P; i += TS) th[i] = init[i];
    } else {
        const double* cr = a.carry + chain * (P + SCCODA_NCARRY);
        for (int i = tid; i < P; i += TS) th[i] = (T)cr[i];
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
// Leapfrog epilogue fused into the gradient's back-propagation: for every parameter i (called once, by its owner)
//   g_work[i] = grad;  p[i] += f * grad;  if (advance) theta_next[i] = theta_cur[i] + eps * p[i]
template <typename T>
struct LeapfrogEpi {
    T* gw;
    T* mom;
    const T* cur;
    T* nxt;
    T f, eps;
    bool kick, advance;
    __device__ __forceinline__ void operator()(int i, T gr) const {
        gw[i] = gr;
        if (kick) {
            const T m = fma(f, gr, mom[i]);
            mom[i] = m;
            if (advance) nxt[i] = fma(eps, m, cur[i]);
        }
    }
};

template <typename T, int W, bool GR>
__global__ void SCCODA_BOUNDS(W) hmc_kernel(const __grid_constant__ KArgs a) {
    DataTile<T> dt;
    ChainScratch cs;
    Group<W> g;
    int b, c;
    if (!prologue<T, W>(a, 6, dt, cs, g, b, c)) return;
    constexpr int TS = 32 * W;
    const int P = a.dm.P, tid = g.tid;
    const uint32_t vs = vec_stride_bytes<T>(P);
    // state vectors (shared-memory byte offsets); the roles rotate on acceptance
    uint32_t o_thc = cs.vec;            // current state
    uint32_t o_gc = cs.vec + vs;        // its gradient
    uint32_t o_a = cs.vec + 2 * vs;     // leapfrog positions (ping-pong)
    uint32_t o_b = cs.vec + 3 * vs;
    uint32_t o_gw = cs.vec + 4 * vs;    // gradient at the latest leapfrog position
    T* const mom = sp<T>(cs.vec + 5 * vs);   // momentum
    const size_t chain = (size_t)b * a.C + c;
    const RngKey key{(uint32_t)(a.seed & 0xffffffffu), (uint32_t)(a.seed >> 32), a.chain_id0 + (uint32_t)chain};

    Adapt<T> ad;
    load_chain_state<T>(a, chain, sp<T>(o_a), ad, tid, TS);
    const int L = a.num_leapfrog;
    const double log_target = log(a.target_accept);
    T* draws = static_cast<T*>(a.draws) + chain * (size_t)a.num_results * P;
    T* stats = static_cast<T*>(a.stats) + chain * (size_t)a.num_results * SCCODA_NSTAT;

    // Transition t = step_begin - 1 is a pseudo-step that only evaluates the starting point (value + gradient)
    // and "accepts" it, so that the hot loop contains exactly one dm_grad and one dm_value call site.
    double lp_cur = 0.0;
    for (int t = a.step_begin - 1; t < a.step_end; ++t) {
        const bool boot = t < a.step_begin;
        const T eps = ad.eps;
        T kin = T(0);
        if (!boot) {
            // momentum draw, first half kick and first drift: q1 = q0 + eps (p + eps/2 grad(q0))
            const T* thc = sp<T>(o_thc);
            const T* gc = sp<T>(o_gc);
            T* qa = sp<T>(o_a);
#pragma unroll 1
            for (int i = tid; i < P; i += TS) {
                const T p = frozen_param(a.dm, i) ? T(0) : rng_normal<T>(key, (uint32_t)i, (uint32_t)t);
                kin += p * p;
                const T m = fma(T(0.5) * eps, gc[i], p);
                mom[i] = m;
                qa[i] = fma(eps, m, thc[i]);
            }
        }
        g.sync();
        uint32_t cur = o_a, nxt = o_b;
        const int nl = boot ? 1 : L;
#pragma unroll 1
        for (int l = 0; l < nl; ++l) {
            const bool last = (l == nl - 1);
            LeapfrogEpi<T> epi{sp<T>(o_gw), mom, sp<T>(cur), sp<T>(nxt), last ? T(0.5) * eps : eps, eps, !boot, !last};
            grad_of<T, W, GR>(a.dm, dt, cs, cur, g, epi);
            if (!last) {
                const uint32_t tmp = cur; cur = nxt; nxt = tmp;
            }
        }
        const double lp_new = value_of<T, W, GR>(a.dm, dt, cs, cur, g);
        bool acc = true;
        double lar = 0.0;
        if (!boot) {
#pragma unroll 1
            for (int i = tid; i < P; i += TS) kin -= mom[i] * mom[i];
            const double dkin = g.sum((double)kin);
            // [TFP] log_accept_ratio = safe_sum([logp', -logp, kinetic - kinetic']); non-finite -> -inf
            lar = lp_new - lp_cur + 0.5 * dkin;
            if (!isfinite(lar)) lar = -CUDART_INF;
            const double u = (double)rng_uniform<T>(key, 0u, (uint32_t)t, STREAM_ACCEPT);
            acc = log(u) < lar;
        }
        if (acc) {
            // the proposal buffer becomes the current state; the old current buffer joins the ping-pong pair
            const uint32_t old = o_thc;
            o_thc = cur;
            if (cur == o_a) o_a = old; else o_b = old;
            const uint32_t tmp = o_gc; o_gc = o_gw; o_gw = tmp;
            lp_cur = lp_new;
        }
        if (boot) continue;
        T traced_eps = eps;
        if (a.method == SCCODA_METHOD_HMC) {
            ad.simple(lar, t, a.num_adapt, log_target);
        } else {
            ad.dual(fmin(lar, 0.0), a.num_adapt, (T)a.target_accept);
            traced_eps = t_exp(ad.log_avg);  // scCODA_model.py:408,419
        }
        const int r = t - a.num_burnin;
        if (r >= 0 && a.store_draws) {
            T* dr = draws + (size_t)r * P;
            const T* thc = sp<T>(o_thc);
#pragma unroll 1
            for (int i = tid; i < P; i += TS) dr[i] = thc[i];
            if (tid == 0)
                write_stats<T>(stats + (size_t)r * SCCODA_NSTAT, lp_cur, lar, acc, traced_eps, lar < -1000.0, L,
                               CUDART_NAN, false);
        }
    }
    store_chain_state<T>(a, chain, sp<T>(o_thc), ad, lp_cur, tid, TS);
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
// The scalar bookkeeping of a NUTS leaf (subtree weight, multinomial acceptance, accept statistic) in the arithmetic
// type: the fp32 build evaluates the transcendental functions in fp32 (the values stay doubles; the fp64 versions
// were ~300 instructions per leaf, executed by every warp of the chain), the fp64 build is unchanged.
template <typename T>
__device__ __forceinline__ double nuts_logaddexp(double x, double y) {
    if constexpr (sizeof(T) == 8) {
        return logaddexp_d(x, y);
    } else {
        if (x == -CUDART_INF) return y;
        if (y == -CUDART_INF) return x;
        return fmax(x, y) + (double)log1pf(expf(-fabsf((float)(x - y))));
    }
}
template <typename T>
__device__ __forceinline__ double nuts_log1m(double u) {   // log(1 - u), u a 24-bit uniform (exact in fp32)
    if constexpr (sizeof(T) == 8) return log1p(-u);
    else return (double)log1pf(-(float)u);
}
template <typename T>
__device__ __forceinline__ double nuts_exp_neg(double d) {  // exp(min(d, 0))
    if constexpr (sizeof(T) == 8) return exp(fmin(d, 0.0));
    else return (double)expf(fminf((float)d, 0.0f));
}

template <typename T, int W, bool GR>
__global__ void SCCODA_BOUNDS(W) nuts_kernel(const __grid_constant__ KArgs a) {
    DataTile<T> dt;
    ChainScratch cs;
    Group<W> g;
    int b, c;
    const int maxd = a.max_depth;
    if (!prologue<T, W>(a, 15 + 2 * (maxd + 1), dt, cs, g, b, c)) return;
    constexpr int TS = 32 * W;
    const int P = a.dm.P, tid = g.tid;
    const uint32_t vsb = vec_stride_bytes<T>(P);
    const size_t vs = vsb / sizeof(T);
    T* const vec = sp<T>(cs.vec);
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
    const uint32_t o_q = cs.vec + 12 * vsb;
    const size_t chain = (size_t)b * a.C + c;
    const RngKey key{(uint32_t)(a.seed & 0xffffffffu), (uint32_t)(a.seed >> 32), a.chain_id0 + (uint32_t)chain};

    Adapt<T> ad;
    load_chain_state<T>(a, chain, q, ad, tid, TS);

    T* draws = static_cast<T*>(a.draws) + chain * (size_t)a.num_results * P;
    T* stats = static_cast<T*>(a.stats) + chain * (size_t)a.num_results * SCCODA_NSTAT;

    // t = step_begin - 1: pseudo-transition that evaluates the starting point through the same (single) call
    // site of dm_grad / dm_value as the leaves of the trees.
    double lp_cur = 0.0;
    for (int t = a.step_begin - 1; t < a.step_end; ++t) {
        const bool boot = t < a.step_begin;
        const T eps = ad.eps;
        double e0 = 0.0;
        if (!boot) {
            T kin = T(0);
#pragma unroll 1
            for (int i = tid; i < P; i += TS) {
                const T pm = frozen_param(a.dm, i) ? T(0) : rng_normal<T>(key, (uint32_t)i, (uint32_t)t);
                kin += pm * pm;
                lq[i] = cq[i]; lpm[i] = pm; lg[i] = cg[i];
                rq[i] = cq[i]; rpm[i] = pm; rg[i] = cg[i];
                rho[i] = pm;
            }
            e0 = lp_cur - 0.5 * g.sum((double)kin);
        }
        double l_lp = lp_cur, r_lp = lp_cur, c_lp = lp_cur, c_energy = e0, s_lp = lp_cur, s_en = e0;
        double wsum = 0.0, accept_sum = 0.0;
        int leapfrogs = 0, depth = 0;
        bool is_acc = false, cont = true, not_div = true;
        uint32_t leaf_serial = 0;
        const int depth_limit = boot ? 1 : maxd;
        while (depth < depth_limit && cont) {
            const bool forward = boot ? true : (rng_bit(key, (uint32_t)depth, (uint32_t)t, STREAM_NUTS_DIR) != 0u);
            T* eq = forward ? rq : lq;
            T* ep = forward ? rpm : lpm;
            T* eg = forward ? rg : lg;
            const T se = forward ? eps : -eps;
            double lqv = forward ? r_lp : l_lp;
            if (!boot) {
#pragma unroll 1
                for (int i = tid; i < P; i += TS) {
                    q[i] = eq[i]; p[i] = ep[i]; gq[i] = eg[i];
                    rho_s[i] = T(0);
                }
            }
            const unsigned n_leaves = 1u << depth;
            double w_s = -CUDART_INF;
            bool alive = true;
            for (unsigned leaf = 0; leaf < n_leaves && alive; ++leaf) {
                if (!boot) {
#pragma unroll 1
                    for (int i = tid; i < P; i += TS) {
                        const T ph = fma(T(0.5) * se, gq[i], p[i]);
                        p[i] = ph;
                        q[i] = fma(se, ph, q[i]);
                    }
                }
                g.sync();
                lqv = grad_value_of<T, W, GR>(a.dm, dt, cs, o_q, g, StoreGrad<T>{gq});
                if (boot) break;
                T kk = T(0);
#pragma unroll 1
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
#pragma unroll 1
                    for (int i = tid; i < P; i += TS) { cp[i] = p[i]; cr[i] = rho_s[i]; }
                } else {
                    const int n_chk = __ffs(~leaf) - 1;  // trailing ones
                    for (int s = idx_max; s > idx_max - n_chk && no_uturn; --s) {
                        const T* cp = ck_p + (size_t)s * vs;
                        const T* cr = ck_rho + (size_t)s * vs;
                        T d1 = T(0), d2 = T(0);
#pragma unroll 1
                        for (int i = tid; i < P; i += TS) {
                            const T sp = rho_s[i] - cr[i] + cp[i];
                            d1 = fma(sp, cp[i], d1);
                            d2 = fma(sp, p[i], d2);
                        }
                        double dd[2] = {(double)d1, (double)d2};
                        g.sumv(dd);
                        if (!(dd[0] >= 0.0 && dd[1] >= 0.0)) no_uturn = false;
                    }
                }
                double energy = lqv - 0.5 * g.sum((double)kk);
                if (isnan(energy)) energy = -CUDART_INF;
                const double delta = energy - e0;
                const bool divergent = !(-delta < 1000.0);
                const double w_new = nuts_logaddexp<T>(w_s, delta);
                const double u = (double)rng_uniform<T>(key, leaf_serial++, (uint32_t)t, STREAM_NUTS_LEAF);
                const double thresh = (w_new > -CUDART_INF) ? delta - w_new : CUDART_NAN;
                if (nuts_log1m<T>(u) <= thresh) {
#pragma unroll 1
                    for (int i = tid; i < P; i += TS) { sq[i] = q[i]; sg[i] = gq[i]; }
                    s_lp = lqv; s_en = energy;
                }
                w_s = w_new;
                accept_sum += nuts_exp_neg<T>(delta);
                if (divergent) not_div = false;
                alive = (!divergent) && no_uturn;
            }
            if (boot) {
                // starting point: becomes the current state
#pragma unroll 1
                for (int i = tid; i < P; i += TS) { cq[i] = q[i]; cg[i] = gq[i]; }
                c_lp = lqv;
                break;
            }
            const double tree_w = alive ? w_s : -CUDART_INF;
            double thresh = tree_w - wsum;
            if (isnan(thresh)) thresh = 0.0;
            const double u = (double)rng_uniform<T>(key, (uint32_t)depth, (uint32_t)t, STREAM_NUTS_TOP);
            if ((nuts_log1m<T>(u) <= thresh) && alive) {
#pragma unroll 1
                for (int i = tid; i < P; i += TS) { cq[i] = sq[i]; cg[i] = sg[i]; }
                c_lp = s_lp; c_energy = s_en; is_acc = true;
            }
            wsum = nuts_logaddexp<T>(wsum, tree_w);
            T d1 = T(0), d2 = T(0);
#pragma unroll 1
            for (int i = tid; i < P; i += TS) {
                eq[i] = q[i]; ep[i] = p[i]; eg[i] = gq[i];
                const T r = rho[i] + rho_s[i];
                rho[i] = r;
                d1 = fma(r, lpm[i], d1);   // ep aliases lpm or rpm: already updated for this thread's i
                d2 = fma(r, rpm[i], d2);
            }
            if (forward) r_lp = lqv; else l_lp = lqv;
            double dd[2] = {(double)d1, (double)d2};
            g.sumv(dd);
            cont = alive && (dd[0] >= 0.0) && (dd[1] >= 0.0);
            depth += 1;
        }
        lp_cur = c_lp;
        if (boot) continue;
        const double lar = accept_sum > 0.0 ? log(accept_sum / (double)leapfrogs) : -CUDART_INF;
        ad.dual(fmin(lar, 0.0), a.num_adapt, (T)a.target_accept);
        const int r = t - a.num_burnin;
        if (r >= 0 && a.store_draws) {
            T* dr = draws + (size_t)r * P;
#pragma unroll 1
            for (int i = tid; i < P; i += TS) dr[i] = cq[i];
            if (tid == 0)
                write_stats<T>(stats + (size_t)r * SCCODA_NSTAT, lp_cur, lar, is_acc, eps, !not_div, leapfrogs,
                               c_energy, cont);
        }
    }
    store_chain_state<T>(a, chain, cq, ad, lp_cur, tid, TS);
}

// ------------------------------------------------------------------------------------------------
// launch helpers (one dtype per translation unit)
// ------------------------------------------------------------------------------------------------
template <typename T, int W, bool GR>
static cudaError_t launch_one(int kind, const KArgs& a, int grid, int block, size_t smem, cudaStream_t st) {
    cudaError_t e;
    if (kind == KIND_LOGP) {
        e = cudaFuncSetAttribute(logp_grad_kernel<T, W, GR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        logp_grad_kernel<T, W, GR><<<grid, block, smem, st>>>(a);
    } else if (kind == KIND_HMC) {
        e = cudaFuncSetAttribute(hmc_kernel<T, W, GR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        hmc_kernel<T, W, GR><<<grid, block, smem, st>>>(a);
    } else {
        e = cudaFuncSetAttribute(nuts_kernel<T, W, GR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        nuts_kernel<T, W, GR><<<grid, block, smem, st>>>(a);
    }
    return cudaGetLastError();
}

template <typename T, bool GR>
static cudaError_t launch_w(int kind, int W, const KArgs& a, int grid, int block, size_t smem, cudaStream_t st) {
    switch (W) {
        case 1: return launch_one<T, 1, GR>(kind, a, grid, block, smem, st);
        case 2: return launch_one<T, 2, GR>(kind, a, grid, block, smem, st);
        case 4: return launch_one<T, 4, GR>(kind, a, grid, block, smem, st);
        case 8: return launch_one<T, 8, GR>(kind, a, grid, block, smem, st);
        case 16: return launch_one<T, 16, GR>(kind, a, grid, block, smem, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace sccoda
