// K2cc — HMC for case/control designs, the shape scCODA is mostly run on and the one the benchmark sweep uses:
// one covariate (D == 1), two covariate groups (U <= 2), K <= 32 cell types, fp32, one warp per chain.
//
// Same algorithm, same Philox streams and same arithmetic per parameter as hmc_kernel (kernels_impl.cuh, which
// follows sccoda/model/scCODA_model.py:276-315, :382-424 and the TFP semantics of SURVEY §8c); what changes is
// where things live.  Lane k owns cell type k for the whole run: alpha_k, b_offset_k, ind_raw_k, their momenta
// and gradients are REGISTERS (sigma_d and its momentum are replicated in every lane), so the leapfrog update is
// a handful of FFMAs without shared-memory traffic, the spike-and-slab effect, both concentrations and both
// pair records of the cell type never leave the lane, and the chain rule needs one warp reduction (sigma_d).
// Shared memory holds the staged dataset, the pair records the item pass reads, and the item results.
#pragma once
#include "kernels_impl.cuh"

namespace sccoda {

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

// The four parameters a lane owns: alpha_k, b_offset_k, ind_raw_k (zero on the reference lane and on lanes >= K)
// and sigma_d (uniform).
struct CcVec {
    float a, b, i, s;
};

struct CcLane {
    int k, j;            // cell type, its index among the non-reference cell types
    bool has, active;    // k < K; k < K and k != ref
    float x0, x1;        // the covariate value of the two groups
};

// Gradient of the log-density at q, pair / item layout (grouped.cuh).  Leaves the pair records
// (concentrations included) in shared memory for dm_value.
__device__ __forceinline__ CcVec cc_grad(const Dims& dm, const DataTile<float>& dt, const ChainScratch& cs,
                                         const CcLane& ln, const CcVec& q, int lane) {
    constexpr int R = SCCODA_ITEM_R;
    const int K = dm.K, tot0 = dt.tot0;
    PairRec<float>* prec = sp<PairRec<float>>(cs.cp);
    float* A = sp<float>(cs.G);
    const float* pcoef = sp<float>(dt.pcoef);
    const float* iy = sp<float>(dt.iy);
    const uint16_t* ipair = sp<uint16_t>(dt.ipair);
    const uint16_t* ipos = sp<uint16_t>(dt.ipos);

    // spike-and-slab effect of the lane's cell type                                             (:724-734)
    const float id = ln.active ? sigmoid_stable(50.0f * q.i) : 0.0f;
    const float bet = id * q.s * q.b;
    // concentrations and pair records of the two groups                                        (:743)
    const int kc = ln.has ? ln.k : 0;
    const float c0 = exp_acc(fmaf(ln.x0, bet, q.a)), c1 = exp_acc(fmaf(ln.x1, bet, q.a));
    const PairRec<float> r0 = make_rec(c0, pcoef + 6 * kc), r1 = make_rec(c1, pcoef + 6 * (K + kc));
    if (ln.has) {
        prec[kc] = r0;
        prec[K + kc] = r1;
    }
    // total pairs: S_u = sum_k c_uk; lanes 0..15 hold S_0, lanes 16..31 hold S_1
    const float S = warp_sum2(ln.has ? c0 : 0.0f, ln.has ? c1 : 0.0f, lane);
    if ((lane & 15) == 0) {
        const int pt = tot0 + (lane >> 4);
        prec[pt] = make_rec(S, pcoef + 6 * pt);
    }
    __syncwarp();
    // items and odd groups, each into its own result slot
    {
        const int NIs = dt.NImax;
#pragma unroll 1
        for (int i = lane; i < dt.NI; i += 32) {
            const PairRec<float> r = prec[ipair[i]];
            float y[R];
#pragma unroll
            for (int jj = 0; jj < R; ++jj) y[jj] = iy[jj * NIs + i];
            A[ipos[i]] = item_eval(r, y);
        }
#pragma unroll 1
        for (int qd = lane; qd < dt.NG; qd += 32) {
            const PairRec<float> r = prec[dt.opair[qd]];
            const uint32_t od = dt.odesc[qd];
            const float* oy = dt.oy + (od & 0xffffu);
            float s = 0.0f;
#pragma unroll 1
            for (int t = 0; t < (int)(od >> 16); ++t) s += odd_eval(r, oy[t]);
            A[dt.opos[qd]] = s;
        }
    }
    __syncwarp();
    // d logp / d log c_uk = c_uk (Br_uk - Br_total(u)), chain rule
    const int NF = dt.NF;   // 2 or 4 result slots per pair
    const float bt0 = pair_bracket<float>(prec[tot0].rat, A + NF * tot0, NF);
    const float bt1 = pair_bracket<float>(prec[tot0 + 1].rat, A + NF * (tot0 + 1), NF);
    const float b0 = pair_bracket<float>(r0.rat, A + NF * kc, NF);
    const float b1 = pair_bracket<float>(r1.rat, A + NF * (K + kc), NF);
    const float G0 = ln.has ? c0 * (b0 - bt0) : 0.0f, G1 = ln.has ? c1 * (b1 - bt1) : 0.0f;
    CcVec g;
    g.a = fmaf(q.a, -0.04f, G0 + G1);                                      // d/d alpha_k   (0 on lanes >= K)
    const float t = fmaf(ln.x0, G0, ln.x1 * G1) * id;                      // id == 0 on inactive lanes
    g.b = fmaf(t, q.s, -q.b);                                              // d/d b_offset_k
    g.i = fmaf(t * q.s * q.b, 50.0f * (1.0f - id), -q.i);                  // d/d ind_raw_k
    // d/d sigma_d: sum_k g_k ind_k b_offset_k - HalfCauchy term (zero prior gradient below 0 [TFP])
    g.s = warp_sum1(t * q.b) - (q.s >= 0.0f ? 2.0f * q.s / (1.0f + q.s * q.s) : 0.0f);
    return g;
}

__global__ void __launch_bounds__(128, 7) hmc_cc_kernel(const __grid_constant__ KArgs a) {
    using T = float;
    DataTile<T> dt;
    ChainScratch cs;
    Group<1> g;
    int b, c;
    if (!prologue<T, 1>(a, 1, dt, cs, g, b, c)) return;
    const int lane = g.tid;
    const int K = a.dm.K, K1 = a.dm.K1, P = a.dm.P;
    T* const vec = sp<T>(cs.vec);                 // one flat state vector: momentum transpose, theta for dm_value
    const size_t chain = (size_t)b * a.C + c;
    const RngKey key{(uint32_t)(a.seed & 0xffffffffu), (uint32_t)(a.seed >> 32), a.chain_id0 + (uint32_t)chain};

    CcLane ln;
    ln.k = lane;
    ln.has = lane < K;
    ln.active = ln.has && lane != dt.ref;
    ln.j = lane - (lane > dt.ref);
    ln.x0 = sp<T>(dt.xu)[0];
    ln.x1 = dt.U > 1 ? sp<T>(dt.xu)[1] : T(0);
    // flat indices of the lane's parameters (order of scCODA_model.py:692-693 with D == 1)
    const int ix_b = 1 + ln.j, ix_i = 1 + K1 + ln.j, ix_a = 1 + 2 * K1 + ln.k;

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
                const T p = rng_normal<T>(key, (uint32_t)i, (uint32_t)t);
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
        CcVec gq;
#pragma unroll 1
        for (int l = 0; l < nl; ++l) {
            gq = cc_grad(a.dm, dt, cs, ln, q, lane);
            if (!boot) {
                const bool last = (l == nl - 1);
                const T f = last ? T(0.5) * eps : eps;
                m.a = fmaf(f, gq.a, m.a); m.b = fmaf(f, gq.b, m.b); m.i = fmaf(f, gq.i, m.i); m.s = fmaf(f, gq.s, m.s);
                if (!last) {
                    q.a = fmaf(eps, m.a, q.a); q.b = fmaf(eps, m.b, q.b); q.i = fmaf(eps, m.i, q.i); q.s = fmaf(eps, m.s, q.s);
                }
            }
        }
        // value at the end point (dm_value reads theta from shared memory and the concentrations cc_grad left)
        __syncwarp();
        scatter(vec, q);
        __syncwarp();
        const double lp_new = dm_value<T, 1, PairRec<T>>(a.dm, dt, cs, cs.vec, g);
        bool acc = true;
        double lar = 0.0;
        if (!boot) {
            // kinetic energies: the momenta of sigma_d are replicated, count them once
            kin -= fmaf(m.a, m.a, fmaf(m.b, m.b, m.i * m.i));
            if (lane == 0) kin -= m.s * m.s;
            const double dkin = g.sum((double)kin);
            // [TFP] log_accept_ratio = safe_sum([logp', -logp, kinetic - kinetic']); non-finite -> -inf
            lar = lp_new - lp_cur + 0.5 * dkin;
            if (!isfinite(lar)) lar = -CUDART_INF;
            const double u = (double)rng_uniform<T>(key, 0u, (uint32_t)t, STREAM_ACCEPT);
            acc = log(u) < lar;
        }
        if (acc) {
            thc = q;
            gc = gq;
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
}

}  // namespace sccoda
