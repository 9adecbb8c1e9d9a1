// K2mg — HMC on the lane-per-cell-type layout of hmc_cc.cuh for designs with MORE than two covariate groups:
// up to 32 groups (distinct rows of the design matrix: a few conditions, 2x2 ... 2x2x2x2 factorials, a continuous
// covariate on a few dozen samples), D <= 4
// covariates, fp32, one warp per chain; K <= 31 cell types with one cell type per lane (NS = 1), K <= 62 with two
// (NS = 2: lane l owns the columns l and 32 + l; the row totals are column K).
//
// Same algorithm, Philox streams and per-parameter arithmetic as hmc_kernel (kernels_impl.cuh; reference:
// sccoda/model/scCODA_model.py:276-315, :382-424) and the same pair / item formulation of the gradient as
// hmc_cc_kernel.  Lane k owns cell type k (lane K the row totals): alpha_k and, per covariate d, b_offset_dk,
// ind_raw_dk with their momenta and gradients are registers; sigma_d is replicated.  The covariate groups are
// walked two at a time on packed fp32 pairs (FFMA2); group pair `up` covers groups 2 up and 2 up + 1, an odd
// group count leaves the last slot empty (no rows, no items, zero coefficients -> contributes nothing).
// The case/control kernel stays the special case D == 1, two groups (everything of a lane fits 72 registers there).
#pragma once
#define SCCODA_CC_HELPERS_ONLY
#include "hmc_cc.cuh"

namespace sccoda {

struct MgTile {
    uint32_t cf, yv, ov, om, rows, el, xg, cnt, bnd;   // shared-memory byte offsets (MgPlan)
    uint32_t yv_stride, ov_stride;                 // bytes per (pair of groups, column slot)
    int UP, UPmax;                                 // pairs of groups of THIS dataset / of the fullest dataset of the batch
    int N, NK, nbig;
    float c_max;                                   // rejection limit of the concentrations (Dims::conc_max)
    double lconst;
};

// (group u, column) of pair index p = u*K + k (cell types) or Umax*K + u (row totals -> column K)
__device__ __forceinline__ void mg_pair_to_col(int p, int K, int Umax, int& u, int& col) {
    if (p >= Umax * K) {
        u = p - Umax * K;
        col = K;
    } else {
        u = p / K;
        col = p - u * K;
    }
}

// keeps a shared-memory offset in a register: without it the compiler rebuilds the offsets from the launch
// arguments (a few dozen integer instructions) wherever register pressure is high
__device__ __forceinline__ uint32_t pin_u32(uint32_t v) {
    asm volatile("mov.u32 %0, %0;" : "+r"(v));
    return v;
}

// Staging: the packed dataset (include/sccoda_b200.h) re-laid lane-major, one block per pair of groups
__device__ __forceinline__ void mg_stage(const KArgs& a, int b, const MgPlan& pl, int UPmax, int NS, MgTile& t) {
    using T = float;
    const int N = a.dm.N, K = a.dm.K, D = a.dm.D, NK = a.dm.NK, Umax = a.Umax, NP = Umax * (K + 1);
    const int NT = a.NT, NOT = a.NOT;
    t.cf = pin_u32((uint32_t)pl.cf); t.yv = pin_u32((uint32_t)pl.yv); t.ov = pin_u32((uint32_t)pl.ov);
    t.om = pin_u32((uint32_t)pl.om);
    t.rows = pin_u32((uint32_t)pl.rows); t.el = pin_u32((uint32_t)pl.el); t.xg = pin_u32((uint32_t)pl.xg);
    t.cnt = pin_u32((uint32_t)pl.cnt); t.bnd = pin_u32((uint32_t)pl.bounds);
    const int NTC = pl.NTC;
    t.yv_stride = pin_u32((uint32_t)(NTC * kItemR * 32 * 8)); t.ov_stride = pin_u32((uint32_t)(NOT * 32 * 8));
    t.N = N; t.NK = NK;
    t.c_max = (float)a.dm.conc_max;
    t.UP = (a.U[b] + 1) >> 1;
    t.UPmax = UPmax;
    const int CS = 32 * NS, NQ = UPmax * NS;
    // [q]: items, [NQ + q]: odd counts of the fullest pair of (group pair, column slot) q = up * NS + s,
    // [2 NQ]: elements with a count >= 6
    int* bounds = sp<int>((uint32_t)pl.bounds);
    for (int i = threadIdx.x; i < 2 * NQ + 1; i += blockDim.x) bounds[i] = 0;
    const int* rstart = a.rstart + (size_t)b * (Umax + 1);
    float* cnt = sp<float>((uint32_t)pl.cnt);
    if (threadIdx.x < 2 * UPmax) {
        const int u = threadIdx.x;
        cnt[u] = u < Umax ? (float)(rstart[u + 1] - rstart[u]) : 0.0f;
    }
    // covariate values of the groups: xg[up][d] = (x of group 2 up, x of group 2 up + 1)
    float* xg = sp<float>((uint32_t)pl.xg);
    const T* xu = static_cast<const T*>(a.xu) + (size_t)b * Umax * D;
    for (int i = threadIdx.x; i < UPmax * D * 2; i += blockDim.x) {
        const int h = i & 1, r = i >> 1, up = r / D, d = r - up * D, u = 2 * up + h;
        xg[i] = u < Umax ? xu[u * D + d] : 0.0f;
    }
    float* cf = sp<float>((uint32_t)pl.cf);
    float* yv = sp<float>((uint32_t)pl.yv);
    float* ov = sp<float>((uint32_t)pl.ov);
    float* om = sp<float>((uint32_t)pl.om);
    int* fill = sp<int>((uint32_t)pl.fill);
    for (int i = threadIdx.x; i < NQ * 64; i += blockDim.x) fill[i] = 0;
    const T* pcoef = static_cast<const T*>(a.pcoef) + (size_t)b * NP * 6;
    // rational coefficients: cf[q][j][lane] = (pair (2 up, col), pair (2 up + 1, col)), col = 32 s + lane; col K = row totals
    for (int i = threadIdx.x; i < NQ * 6 * 32 * 2; i += blockDim.x) {
        const int h = i & 1, lane = (i >> 1) & 31, r = i >> 6, q = r / 6, j = r - q * 6, up = q / NS, col = (q - up * NS) * 32 + lane;
        const int u = 2 * up + h;
        float v = 0.0f;
        if (col <= K && u < Umax) v = pcoef[(col < K ? u * K + col : Umax * K + u) * 6 + j];
        cf[i] = v;
    }
    // item counts: padding is the count 6, which contributes psi(c+6) - psi(c+6) = 0; odd counts: padding has the
    // multiplicity 0
    for (int i = threadIdx.x; i < NQ * NTC * kItemR * 64; i += blockDim.x) yv[i] = 6.0f;
    for (int i = threadIdx.x; i < NQ * NOT * 64; i += blockDim.x) {
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
        mg_pair_to_col(p, K, Umax, u, col);
        const int q = (u >> 1) * NS + (col >> 5);
        if (tt < 0 || tt >= NT || col > K || u >= 2 * UPmax) continue;   // never write outside the tile
        yv[(((q * NTC + tt) * kItemR + j) * 32 + (col & 31)) * 2 + (u & 1)] = iy[j * a.NImax + it];
        if (j == 0) {
            atomicMax(&bounds[q], tt + 1);
            atomicMax(&fill[(q * 32 + (col & 31)) * 2 + (u & 1)], tt * kItemR + (int)a.inreal[(size_t)b * a.NImax + it]);
        }
    }
    __syncthreads();
    // odd counts: y + 6 into the next free item slots of the pair, -m sum_{i<6} 1/(c + y + i) per distinct value y
    // (cc_stage in hmc_cc.cuh explains the identity)
    const int NG = a.NG[b];
    const T* oy = static_cast<const T*>(a.oy) + (size_t)b * a.NOmax;
    for (int q = threadIdx.x; q < NG; q += blockDim.x) {
        const uint32_t od = a.odesc[(size_t)b * a.NGmax + q];
        int u, col;
        mg_pair_to_col(a.opair[(size_t)b * a.NGmax + q], K, Umax, u, col);
        if (col > K || u >= 2 * UPmax) continue;
        const int qq = (u >> 1) * NS + (col >> 5), cl = col & 31, h = u & 1;
        const int base = fill[(qq * 32 + cl) * 2 + h];
        int n = (int)(od >> 16);
        if (n > NOT) n = NOT;
        int nd = 0;
        for (int tt = 0; tt < n; ++tt) {
            const float y = oy[(od & 0xffffu) + tt];
            const int slot = base + tt;                          // item * R + position
            if (slot >= NTC * kItemR) break;
            yv[((qq * NTC * kItemR + slot) * 32 + cl) * 2 + h] = y + 6.0f;
            int d = 0;
            while (d < nd && ov[((qq * NOT + d) * 32 + cl) * 2 + h] != y) ++d;
            if (d == nd) {
                ov[((qq * NOT + d) * 32 + cl) * 2 + h] = y;
                ++nd;
            }
            om[((qq * NOT + d) * 32 + cl) * 2 + h] += 1.0f;
        }
        if (n > 0) atomicMax(&bounds[qq], (base + n - 1) / kItemR + 1);
        atomicMax(&bounds[NQ + qq], nd);
    }
    // value pass, rows and elements: as in cc_stage (hmc_cc.cuh), concentration slot = group * CS + column
    float4* rows = sp<float4>((uint32_t)pl.rows);
    const T* ntot = static_cast<const T*>(a.ntot) + (size_t)b * N;
    const int* grp = a.grp + (size_t)b * N;
    double lg_sum = 0.0;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        const double nn = (double)ntot[n];
        const double lg = lgamma(nn);
        const double rc = nn >= 6.0 ? lg - ((nn - 0.5) * log(nn) - nn + 0.91893853320467274178) : lg;
        rows[n] = make_float4((float)nn, (float)rc, __int_as_float(grp[n] * CS + K), 0.0f);
        lg_sum += lg;
    }
    double* red = sp<double>((uint32_t)pl.red);
    __syncthreads();
    for (int o = 16; o > 0; o >>= 1) lg_sum += __shfl_xor_sync(0xffffffffu, lg_sum, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = lg_sum;   // one slot per warp (up to 32): deterministic sum below
    float4* el = sp<float4>((uint32_t)pl.el);
    const T* ey = static_cast<const T*>(a.ey) + (size_t)b * NK;
    const T* eycst = static_cast<const T*>(a.eycst) + (size_t)b * NK;
    const uint32_t* eidx = a.eidx + (size_t)b * NK;
    if (threadIdx.x < 32) {
        // counts >= 6 first; partitioned by warp 0 with ballots, in list order (launch-independent summation order)
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
                const int u = uk / K;
                const int dst = big ? ib + __popc(mb & lt) : is + __popc(ms & lt);
                el[dst] = make_float4(y, eycst[e], __int_as_float(u * CS + (uk - u * K)), 1.0f / y);
            }
            ib += __popc(mb);
            is += __popc(ms);
        }
        if (lane == 0) bounds[2 * NQ] = nbig;
    }
    __syncthreads();
    t.nbig = bounds[2 * NQ];
    double lg_all = 0.0;
    for (int w = 0; w < (int)((blockDim.x + 31) >> 5); ++w) lg_all += red[w];
    t.lconst = a.lconst[b] - lg_all;
}

// The parameters a lane owns: per column slot s (cell type 32 s + lane) alpha and, per covariate, b_offset and
// ind_raw (zero on the reference column and on columns >= K), and sigma_d (uniform).
template <int D, int NS>
struct MgVec {
    float a[NS], b[NS][D], i[NS][D], s[D];
};
template <int D, int NS>
__device__ __forceinline__ MgVec<D, NS> mg_zero() {
    MgVec<D, NS> v;
#pragma unroll
    for (int d = 0; d < D; ++d) v.s[d] = 0.0f;
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        v.a[s] = 0.0f;
#pragma unroll
        for (int d = 0; d < D; ++d) v.b[s][d] = v.i[s][d] = 0.0f;
    }
    return v;
}
// y = f x + y
template <int D, int NS>
__device__ __forceinline__ void mg_axpy(MgVec<D, NS>& y, float f, const MgVec<D, NS>& x) {
#pragma unroll
    for (int d = 0; d < D; ++d) y.s[d] = fmaf(f, x.s[d], y.s[d]);
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        y.a[s] = fmaf(f, x.a[s], y.a[s]);
#pragma unroll
        for (int d = 0; d < D; ++d) {
            y.b[s][d] = fmaf(f, x.b[s][d], y.b[s][d]);
            y.i[s][d] = fmaf(f, x.i[s][d], y.i[s][d]);
        }
    }
}

template <int NS>
struct MgLane {
    int K;                     // the row totals are column K = 32 ks() + kl()
    bool has[NS], active[NS];  // column < K; column < K and column != ref
    __device__ __forceinline__ int ks() const { return NS == 1 ? 0 : K >> 5; }
    __device__ __forceinline__ int kl() const { return NS == 1 ? K : K & 31; }
    __device__ __forceinline__ uint32_t hasmask() const {
        uint32_t m = 0;
#pragma unroll
        for (int s = 0; s < NS; ++s) m |= has[s] ? 1u << s : 0u;
        return m;
    }
};

// Gradient of the log-density at q, pair / item formulation; the concentrations of every group are left
// in the chain's shared-memory slots pc[group * 32 NS + column] (column K: S_u) for the value pass.
template <int D, int NS>
__device__ __forceinline__ MgVec<D, NS> mg_grad(const MgTile& t, const MgLane<NS>& ln, const MgVec<D, NS>& q, uint32_t pc_off,
                                                int lane) {
    float id[NS][D], bet[NS][D], tg[NS][D], ga[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        ga[s] = 0.0f;
#pragma unroll
        for (int d = 0; d < D; ++d) {
            id[s][d] = ln.active[s] ? sigmoid_stable(50.0f * q.i[s][d]) : 0.0f;            // (:724-734)
            bet[s][d] = id[s][d] * q.s[d] * q.b[s][d];
            tg[s][d] = 0.0f;
        }
    }
    uint32_t cf_off = t.cf + lane * 8, yv_off = t.yv + lane * 8, ov_off = t.ov + lane * 8, xg_off = t.xg;
    const uint32_t om_delta = t.om - t.ov;
    uint32_t pcw = pc_off + lane * 4;
    const float* cnt = sp<float>(t.cnt);
    const int* bnd = sp<int>(t.bnd);
    const int n_q = t.UPmax * NS;
    __syncwarp();
#pragma unroll 1
    for (int up = 0; up < t.UP; ++up) {
        const f32x2* xg = sp<f32x2>(xg_off);
        const bool empty_hi = cnt[2 * up + 1] == 0.0f;                            // empty slot of an odd group count
        f32x2 c[NS];
        float c0[NS], c1[NS], s0 = 0.0f, s1 = 0.0f;
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            f32x2 lin = dup2(q.a[s]);
#pragma unroll
            for (int d = 0; d < D; ++d) lin = fma2(xg[d], dup2(bet[s][d]), lin);
            c[s] = exp2_acc(lin);                                                 // (:743)
            if (empty_hi) c[s] = pk2(lo2(c[s]), 1.0f);
            c0[s] = ln.has[s] ? lo2(c[s]) : 0.0f;
            c1[s] = ln.has[s] ? hi2(c[s]) : 0.0f;
            s0 += c0[s];
            s1 += c1[s];
        }
        const float S = warp_sum2(s0, s1, lane);                                  // lanes 0..15: S_0, lanes 16..31: S_1
        const float So = __shfl_xor_sync(0xffffffffu, S, 16);
        f32x2 br[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            if (s == ln.ks() && lane == ln.kl()) c[s] = lane < 16 ? pk2(S, So) : pk2(So, S);
            *sp<float>(pcw + s * 128) = lo2(c[s]);
            *sp<float>(pcw + s * 128 + NS * 128) = hi2(c[s]);
            // reference point of the items: x6 = c + 5.5, w6 = 1/x6, a = w6^5, nb = -5 (psi(c+6) - ln x6)
            const f32x2 cm = add2(c[s], dup2(-0.5f));
            const f32x2 w6 = rcp2(add2(c[s], dup2(5.5f)));
            const f32x2 w62 = mul2(w6, w6);
            const f32x2 a5 = mul2(mul2(w62, w62), w6);
            const f32x2 nb = mul2(psi_mid_acc2(w6, dup2(0.0f)), dup2(-(float)SCCODA_ITEM_R));
            // brackets Br = rational sum_i W_i/(c+i) + items + odd counts, both groups at once
            f32x2 b2 = pair_rational2(c[s], sp<f32x2>(cf_off));
            const f32x2* yv = sp<f32x2>(yv_off);
            const int nt = bnd[up * NS + s], n_odd = bnd[n_q + up * NS + s];
#pragma unroll 1
            for (int tt = 0; tt < nt; ++tt, yv += SCCODA_ITEM_R * 32) b2 = add2(b2, item_eval2(cm, a5, nb, yv));
            // odd counts: the shifted values sit in the items above; -m sum_{i<6} 1/(c + y + i) per distinct value
            yv = sp<f32x2>(ov_off);
#pragma unroll 1
            for (int tt = 0; tt < n_odd; ++tt, yv += 32)
                b2 = fma2(*reinterpret_cast<const f32x2*>(reinterpret_cast<const char*>(yv) + om_delta), r6_neg2(add2(c[s], *yv)), b2);
            br[s] = b2;
            cf_off += 6 * 32 * 8; yv_off += t.yv_stride; ov_off += t.ov_stride;
        }
        // d logp / d log c_uk = c_uk (Br_uk - Br_total(u)), chain rule through alpha and the effects
        f32x2 brt = br[0];
#pragma unroll
        for (int s = 1; s < NS; ++s)
            if (s == ln.ks()) brt = br[s];
        const float bt0 = __shfl_sync(0xffffffffu, lo2(brt), ln.kl()), bt1 = __shfl_sync(0xffffffffu, hi2(brt), ln.kl());
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            const float G0 = c0[s] * (lo2(br[s]) - bt0), G1 = c1[s] * (hi2(br[s]) - bt1);   // c0 = c1 = 0 on columns >= K
            ga[s] += G0 + G1;
#pragma unroll
            for (int d = 0; d < D; ++d) {
                const f32x2 xv = xg[d];
                tg[s][d] = fmaf(lo2(xv), G0, fmaf(hi2(xv), G1, tg[s][d]));
            }
        }
        xg_off += D * 8; pcw += NS * 256;
    }
    MgVec<D, NS> g;
    float sg[D];
#pragma unroll
    for (int d = 0; d < D; ++d) sg[d] = 0.0f;
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        g.a[s] = fmaf(q.a[s], -0.04f, ga[s]);                                      // d/d alpha_k
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const float tgi = tg[s][d] * id[s][d];                                 // id == 0 on inactive columns
            g.b[s][d] = fmaf(tgi, q.s[d], -q.b[s][d]);                             // d/d b_offset_dk
            g.i[s][d] = fmaf(tgi * q.s[d] * q.b[s][d], 50.0f * (1.0f - id[s][d]), -q.i[s][d]);   // d/d ind_raw_dk
            sg[d] = fmaf(tgi, q.b[s][d], sg[d]);
        }
    }
#pragma unroll
    for (int d = 0; d < D; ++d) {
        // d/d sigma_d: sum_k g_k ind_k b_offset_k - HalfCauchy term (zero prior gradient below 0 [TFP])
        g.s[d] = warp_sum1(sg[d]) - (q.s[d] >= 0.0f ? 2.0f * q.s[d] * rcp_fast(fmaf(q.s[d], q.s[d], 1.0f)) : 0.0f);
    }
    return g;
}

// lane's share of the parameter priors (:703-741) and the HalfCauchy support flag, for mg_value
template <int D, int NS>
__device__ __forceinline__ float mg_prior(const MgVec<D, NS>& q, int lane, bool& neg_sigma) {
    float quad = 0.0f, slab = 0.0f, cauchy = 0.0f;
    neg_sigma = false;
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        quad = fmaf(q.a[s], q.a[s], quad);
#pragma unroll
        for (int d = 0; d < D; ++d) slab += fmaf(q.b[s][d], q.b[s][d], q.i[s][d] * q.i[s][d]);
    }
#pragma unroll
    for (int d = 0; d < D; ++d) {
        if (q.s[d] >= 0.0f) cauchy += log1pf(q.s[d] * q.s[d]);
        else neg_sigma = true;
    }
    float prior = fmaf(-0.5f, slab, quad * -0.02f);
    if (lane == 0) prior -= cauchy;
    return prior;
}

// Value of the log-density from the concentrations the last gradient evaluation left in pc (same terms, guards and
// precision strategy as cc_value / dm_value); `prior` is the lane's share of the parameter priors, bit s of
// `hasmask` says whether the lane's column slot s holds a cell type.
template <int NS>
static __device__ __noinline__ double mg_value(const MgTile& t, uint32_t pc_off, uint32_t hasmask, float prior, bool neg_sigma,
                                               int lane) {
    constexpr int CS = 32 * NS;
    const float* pc = sp<float>(pc_off);
    __syncwarp();
    double lp = (double)prior;
    // DirichletMultinomial [TFP]: lbeta(c+y) - lbeta(c)                               (:746-751)
    // Guard (numerics): beyond c_max the lgamma differences have lost all precision => NaN => rejected.
    const float c_max = t.c_max, c_joint = SCCODA_C_JOINT;
    bool need_fix = false, bad = false;
    {
        const float* cnt = sp<float>(t.cnt);
        float acc = 0.0f;
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            if (!((hasmask >> s) & 1u)) continue;
#pragma unroll 1
            for (int u = 0; u < 2 * t.UP; ++u) {
                const float cu = pc[u * CS + s * 32 + lane];
                if (cu < c_joint) acc = fmaf(-cnt[u], lgamma_pos(cu), acc); else need_fix = true;
                bad = bad || !(cu <= c_max);
            }
        }
        lp += (double)acc;
    }
    if (bad) lp = CUDART_NAN;
    {
        // rows: -[lgamma(S_u + n) - lgamma(S_u) - lgamma(n)], joint form for S_u >= 6 (the usual case)
        const float4* rows = sp<float4>(t.rows);
        float acc = 0.0f;
#pragma unroll 1
        for (int n = lane; n < t.N; n += 32) {
            const float4 r = rows[n];
            const float S = pc[__float_as_int(r.z)];
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
            // counts >= 6 two at a time, the form chosen per element (lgamma_mixed_big2); the rest one by one
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
    if (neg_sigma) lp = -CUDART_INF;   // HalfCauchy support [TFP]
    return lp;
}

// Per-lane bookkeeping shared by the HMC and the NUTS kernel: which columns the lane owns and where its parameters
// sit in the flat state vector (order of scCODA_model.py:692-693): sigma_d[D] | b_offset[D][K-1] | ind_raw[D][K-1] |
// alpha[K]
template <int D, int NS>
struct MgMap {
    MgLane<NS> ln;
    int ix_a[NS], ix_b[NS], ix_i[NS], K1;

    __device__ __forceinline__ void init(int K, int ref, int lane) {
        K1 = K - 1;
        ln.K = K;
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            const int col = s * 32 + lane;
            ln.has[s] = col < K;
            ln.active[s] = ln.has[s] && col != ref;
            const int jj = col - (col > ref);
            ix_b[s] = D + jj;
            ix_i[s] = D + D * K1 + jj;
            ix_a[s] = D + 2 * D * K1 + col;
        }
    }
    __device__ __forceinline__ MgVec<D, NS> gather(const float* v) const {   // flat shared-memory vector -> lane registers
        MgVec<D, NS> r;
#pragma unroll
        for (int d = 0; d < D; ++d) r.s[d] = v[d];
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            r.a[s] = ln.has[s] ? v[ix_a[s]] : 0.0f;
#pragma unroll
            for (int d = 0; d < D; ++d) {
                r.b[s][d] = ln.active[s] ? v[ix_b[s] + d * K1] : 0.0f;
                r.i[s][d] = ln.active[s] ? v[ix_i[s] + d * K1] : 0.0f;
            }
        }
        return r;
    }
    __device__ __forceinline__ void scatter(float* v, const MgVec<D, NS>& r, int lane) const {   // lane registers -> flat vector
#pragma unroll
        for (int d = 0; d < D; ++d)
            if (lane == 0) v[d] = r.s[d];
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            if (ln.has[s]) v[ix_a[s]] = r.a[s];
            if (ln.active[s]) {
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    v[ix_b[s] + d * K1] = r.b[s][d];
                    v[ix_i[s] + d * K1] = r.i[s][d];
                }
            }
        }
    }
};

// per-lane part of a dot product over the flat parameter vector (sigma_d is replicated: lane 0 counts it)
template <int D, int NS>
__device__ __forceinline__ float mg_dot(const MgVec<D, NS>& x, const MgVec<D, NS>& y, int lane) {
    float v = 0.0f, sv = 0.0f;
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        v = fmaf(x.a[s], y.a[s], v);
#pragma unroll
        for (int d = 0; d < D; ++d) v = fmaf(x.b[s][d], y.b[s][d], fmaf(x.i[s][d], y.i[s][d], v));
    }
#pragma unroll
    for (int d = 0; d < D; ++d) sv = fmaf(x.s[d], y.s[d], sv);
    return lane == 0 ? v + sv : v;
}

// (mg_hmc_ctas / mg_nuts_ctas, the resident CTAs of 128 threads per SM the kernels are compiled for, live in kernels.h)

// ------------------------------------------------------------------------------------------------
// Kernel
// ------------------------------------------------------------------------------------------------
// WIDE: the launch geometry of hmc_cc_wide_kernel (hmc_cc.cuh) — one CTA per SM with a balanced share of all chains of the
// launch, meeting at a barrier every 16 transitions so that the unfair warp schedulers cannot drain the SM early.
template <int D, int NS, bool WIDE>
__device__ __forceinline__ void hmc_mg_body(const KArgs& a) {
    using T = float;
    using V = MgVec<D, NS>;
    const int lane = threadIdx.x & 31, gid = threadIdx.x >> 5;
    const int K = a.dm.K, P = a.dm.P;
    const int UPmax = (a.Umax + 1) >> 1;
    const MgPlan pl = plan_mg(a.dm.N, K, D, UPmax, NS, a.NT, a.NOT, 0);
    MgTile tile;
    int b, c;
    size_t data_bytes = pl.data_total;
    if constexpr (WIDE) {
        const long long n_chains = (long long)a.B * a.C;
        const int lo = (int)(n_chains * blockIdx.x / gridDim.x), hi = (int)(n_chains * (blockIdx.x + 1) / gridDim.x);
        const int chain_g = lo + gid;
        const int b_lo = lo / a.C, b_hi = (hi - 1) / a.C;
        b = chain_g / a.C;
        c = chain_g - b * a.C;
        data_bytes = (size_t)wide_datasets(a.wide, a.C) * pl.data_total;
#pragma unroll 1
        for (int bb = b_lo; bb <= b_hi; ++bb) {
            MgPlan pd = pl;
            const size_t sh = (size_t)(bb - b_lo) * pl.data_total;
            pd.cf += sh; pd.yv += sh; pd.ov += sh; pd.om += sh; pd.rows += sh; pd.el += sh; pd.xg += sh; pd.cnt += sh;
            pd.bounds += sh; pd.red += sh; pd.fill += sh;
            MgTile td;
            mg_stage(a, bb, pd, UPmax, NS, td);
            if (bb == b) tile = td;
        }
        if (chain_g >= hi) return;
    } else {
        const int cpc = blockDim.x >> 5;
        const int ctas_per_ds = (a.C + cpc - 1) / cpc;
        b = blockIdx.x / ctas_per_ds;
        c = (blockIdx.x - b * ctas_per_ds) * cpc + gid;
        mg_stage(a, b, pl, UPmax, NS, tile);
        if (c >= a.C) return;
    }
    const uint32_t chain_off = (uint32_t)(data_bytes + (size_t)gid * pl.chain_total);
    const uint32_t pc_off = pin_u32(chain_off + (uint32_t)pl.pc);
    T* const vec = sp<T>(pin_u32(chain_off + (uint32_t)pl.vec));   // one flat state vector: momentum transpose, carry in / out
    const size_t chain = (size_t)b * a.C + c;
    const RngKey key{(uint32_t)(a.seed & 0xffffffffu), (uint32_t)(a.seed >> 32), a.chain_id0 + (uint32_t)chain};
    MgMap<D, NS> mp;
    mp.init(K, a.ref[b], lane);

    Adapt<T> ad;
    load_chain_state<T>(a, chain, vec, ad, lane, 32);
    __syncwarp();
    V thc = mp.gather(vec);                        // current state
    V gc = mg_zero<D, NS>();                       // its gradient
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
        V q = thc, m = mg_zero<D, NS>();
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
            // first half kick and first drift: q1 = q0 + eps (p + eps/2 grad(q0))
            m = mp.gather(vec);
            mg_axpy<D, NS>(m, T(0.5) * eps, gc);
            mg_axpy<D, NS>(q, eps, m);
        }
        const int nl = boot ? 1 : L;
        V gr;
#pragma unroll 1
        for (int l = 0; l < nl; ++l) {
            gr = mg_grad<D, NS>(tile, mp.ln, q, pc_off, lane);
            if (a.dm.freeze) {                     // SimpleModel: sigma_d and ind_raw stay where they are
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    gr.s[d] = T(0);
#pragma unroll
                    for (int s = 0; s < NS; ++s) gr.i[s][d] = T(0);
                }
            }
            if (!boot) {
                const bool last = (l == nl - 1);
                mg_axpy<D, NS>(m, last ? T(0.5) * eps : eps, gr);
                if (!last) mg_axpy<D, NS>(q, eps, m);
            }
        }
        // value at the end point, from the concentrations of the last gradient evaluation; priors (:703-741), the
        // theta-independent constants live in lconst, parameters a lane does not own are zero
        bool neg_sigma;
        const float prior = mg_prior<D, NS>(q, lane, neg_sigma);
        const double lp_new = mg_value<NS>(tile, pc_off, mp.ln.hasmask(), prior, neg_sigma, lane);
        bool acc = true;
        double lar = 0.0;
        if (!boot) {
            // kinetic energies: sum p0^2 over the flat vector minus the lane's share of the end momenta
            kin -= mg_dot<D, NS>(m, m, lane);
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
            gc = gr;
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
            mp.scatter(draws + (size_t)r * P, thc, lane);
            if (lane == 0)
                write_stats<T>(stats + (size_t)r * SCCODA_NSTAT, lp_cur, lar, acc, traced_eps, lar < -1000.0, L,
                               CUDART_NAN, false);
        }
    }
    __syncwarp();
    mp.scatter(vec, thc, lane);
    __syncwarp();
    store_chain_state<T>(a, chain, vec, ad, lp_cur, lane, 32);
}

template <int D, int NS>
__global__ void __launch_bounds__(128, mg_hmc_ctas(D, NS)) hmc_mg_kernel(const __grid_constant__ KArgs a) {
    hmc_mg_body<D, NS, false>(a);
}
template <int D, int NS>
__global__ void __launch_bounds__(128 * mg_hmc_ctas(D, NS), 1) hmc_mg_wide_kernel(const __grid_constant__ KArgs a) {
    hmc_mg_body<D, NS, true>(a);
}

// K1mg — value and gradient at given states through mg_grad / mg_value (sccoda_logp_grad_as): theta[B,C,P] in a.init,
// gradient to a.draws, value to a.carry[chain].
template <int D, int NS>
__global__ void __launch_bounds__(128) logp_grad_mg_kernel(const __grid_constant__ KArgs a) {
    using T = float;
    using V = MgVec<D, NS>;
    const int lane = threadIdx.x & 31, gid = threadIdx.x >> 5;
    const int cpc = blockDim.x >> 5;
    const int ctas_per_ds = (a.C + cpc - 1) / cpc;
    const int b = blockIdx.x / ctas_per_ds;
    const int c = (blockIdx.x - b * ctas_per_ds) * cpc + gid;
    const int K = a.dm.K, P = a.dm.P;
    const int UPmax = (a.Umax + 1) >> 1;
    const MgPlan pl = plan_mg(a.dm.N, K, D, UPmax, NS, a.NT, a.NOT, 0);
    MgTile tile;
    mg_stage(a, b, pl, UPmax, NS, tile);
    if (c >= a.C) return;
    const uint32_t chain_off = (uint32_t)(pl.data_total + (size_t)gid * pl.chain_total);
    const uint32_t pc_off = pin_u32(chain_off + (uint32_t)pl.pc);
    T* const vec = sp<T>(pin_u32(chain_off + (uint32_t)pl.vec));
    const size_t chain = (size_t)b * a.C + c;
    MgMap<D, NS> mp;
    mp.init(K, a.ref[b], lane);
    const T* theta = static_cast<const T*>(a.init) + chain * P;
    for (int i = lane; i < P; i += 32) vec[i] = theta[i];
    __syncwarp();
    const V q = mp.gather(vec);
    const V gr = mg_grad<D, NS>(tile, mp.ln, q, pc_off, lane);
    bool neg_sigma;
    const float prior = mg_prior<D, NS>(q, lane, neg_sigma);
    const double lp = mg_value<NS>(tile, pc_off, mp.ln.hasmask(), prior, neg_sigma, lane);
    mp.scatter(static_cast<T*>(a.draws) + chain * P, gr, lane);
    if (lane == 0) a.carry[chain] = lp;
}

// ------------------------------------------------------------------------------------------------
// K3mg — NUTS for the same designs: the iterative tree of nuts_kernel / nuts_cc_kernel ([TFP] NoUTurnSampler.one_step as
// restated in SURVEY §8c) on the lane-per-cell-type state.  A stored vector of the tree is mg_vec_quads(D, NS) float4
// per lane (NS (1 + 2 D) + D parameters) in shared memory; the leaf being integrated lives in registers.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double mg_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <int D, int NS>
__device__ __forceinline__ MgVec<D, NS> mg_ld(const float4* slot) {
    constexpr int NQ = mg_vec_quads(D, NS);
    float f[4 * NQ];
#pragma unroll
    for (int k = 0; k < NQ; ++k) {
        const float4 v = slot[32 * k];
        f[4 * k] = v.x; f[4 * k + 1] = v.y; f[4 * k + 2] = v.z; f[4 * k + 3] = v.w;
    }
    MgVec<D, NS> r;
#pragma unroll
    for (int d = 0; d < D; ++d) r.s[d] = f[d];
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        r.a[s] = f[D + s * (1 + 2 * D)];
#pragma unroll
        for (int d = 0; d < D; ++d) {
            r.b[s][d] = f[D + s * (1 + 2 * D) + 1 + d];
            r.i[s][d] = f[D + s * (1 + 2 * D) + 1 + D + d];
        }
    }
    return r;
}
template <int D, int NS>
__device__ __forceinline__ void mg_st(float4* slot, const MgVec<D, NS>& r) {
    constexpr int NQ = mg_vec_quads(D, NS);
    float f[4 * NQ];
#pragma unroll
    for (int k = 0; k < 4 * NQ; ++k) f[k] = 0.0f;
#pragma unroll
    for (int d = 0; d < D; ++d) f[d] = r.s[d];
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        f[D + s * (1 + 2 * D)] = r.a[s];
#pragma unroll
        for (int d = 0; d < D; ++d) {
            f[D + s * (1 + 2 * D) + 1 + d] = r.b[s][d];
            f[D + s * (1 + 2 * D) + 1 + D + d] = r.i[s][d];
        }
    }
#pragma unroll
    for (int k = 0; k < NQ; ++k) slot[32 * k] = make_float4(f[4 * k], f[4 * k + 1], f[4 * k + 2], f[4 * k + 3]);
}

template <int D, int NS>
__global__ void __launch_bounds__(128, mg_nuts_ctas(D, NS)) nuts_mg_kernel(const __grid_constant__ KArgs a) {
    using T = float;
    using V = MgVec<D, NS>;
    constexpr int NQ = mg_vec_quads(D, NS);
    const int lane = threadIdx.x & 31, gid = threadIdx.x >> 5;
    const int cpc = blockDim.x >> 5;
    const int ctas_per_ds = (a.C + cpc - 1) / cpc;
    const int b = blockIdx.x / ctas_per_ds;
    const int c = (blockIdx.x - b * ctas_per_ds) * cpc + gid;
    const int K = a.dm.K, P = a.dm.P, maxd = a.max_depth;
    const int UPmax = (a.Umax + 1) >> 1;
    const MgPlan pl = plan_mg(a.dm.N, K, D, UPmax, NS, a.NT, a.NOT, 11 + 2 * (maxd + 1));
    MgTile tile;
    mg_stage(a, b, pl, UPmax, NS, tile);
    if (c >= a.C) return;
    const uint32_t chain_off = (uint32_t)(pl.data_total + (size_t)gid * pl.chain_total);
    const uint32_t pc_off = pin_u32(chain_off + (uint32_t)pl.pc);
    T* const vec = sp<T>(pin_u32(chain_off + (uint32_t)pl.vec));
    float4* const slots = sp<float4>(pin_u32(chain_off + (uint32_t)pl.slots)) + lane;   // slot k of this lane: slot(k)
    auto slot = [&](int k) { return slots + 32 * NQ * k; };
    enum { S_CQ = 0, S_CG, S_LQ, S_LP, S_LG, S_RQ, S_RP, S_RG, S_SQ, S_SG, S_RHO, S_CKP };
    const int S_CKR = S_CKP + (maxd + 1);
    const size_t chain = (size_t)b * a.C + c;
    const RngKey key{(uint32_t)(a.seed & 0xffffffffu), (uint32_t)(a.seed >> 32), a.chain_id0 + (uint32_t)chain};
    MgMap<D, NS> mp;
    mp.init(K, a.ref[b], lane);

    Adapt<T> ad;
    load_chain_state<T>(a, chain, vec, ad, lane, 32);
    __syncwarp();
    V q = mp.gather(vec), p = mg_zero<D, NS>(), gq = mg_zero<D, NS>(), rho_s = mg_zero<D, NS>();
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
            const V pm = mp.gather(vec), cq = mg_ld<D, NS>(slot(S_CQ)), cg = mg_ld<D, NS>(slot(S_CG));
            mg_st<D, NS>(slot(S_LQ), cq); mg_st<D, NS>(slot(S_LP), pm); mg_st<D, NS>(slot(S_LG), cg);
            mg_st<D, NS>(slot(S_RQ), cq); mg_st<D, NS>(slot(S_RP), pm); mg_st<D, NS>(slot(S_RG), cg);
            mg_st<D, NS>(slot(S_RHO), pm);
            e0 = lp_cur - 0.5 * mg_sum_d((double)kin);
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
                q = mg_ld<D, NS>(slot(e_q));
                p = mg_ld<D, NS>(slot(e_p));
                gq = mg_ld<D, NS>(slot(e_g));
                rho_s = mg_zero<D, NS>();
            }
            const unsigned n_leaves = 1u << depth;
            double w_s = -CUDART_INF;
            bool alive = true;
#pragma unroll 1
            for (unsigned leaf = 0; leaf < n_leaves && alive; ++leaf) {
                if (!boot) {
                    mg_axpy<D, NS>(p, T(0.5) * se, gq);
                    mg_axpy<D, NS>(q, se, p);
                }
                gq = mg_grad<D, NS>(tile, mp.ln, q, pc_off, lane);
                if (a.dm.freeze) {
#pragma unroll
                    for (int d = 0; d < D; ++d) {
                        gq.s[d] = T(0);
#pragma unroll
                        for (int s = 0; s < NS; ++s) gq.i[s][d] = T(0);
                    }
                }
                {
                    bool neg_sigma;
                    const float prior = mg_prior<D, NS>(q, lane, neg_sigma);
                    lqv = mg_value<NS>(tile, pc_off, mp.ln.hasmask(), prior, neg_sigma, lane);
                }
                if (boot) break;
                mg_axpy<D, NS>(p, T(0.5) * se, gq);
                mg_axpy<D, NS>(rho_s, T(1), p);
                const T kk = mg_dot<D, NS>(p, p, lane);
                leapfrogs += 1;
                const int idx_max = __popc(leaf >> 1);
                bool no_uturn = true;
                if ((leaf & 1u) == 0u) {
                    mg_st<D, NS>(slot(S_CKP + idx_max), p);
                    mg_st<D, NS>(slot(S_CKR + idx_max), rho_s);
                } else {
                    const int n_chk = __ffs(~leaf) - 1;  // trailing ones
#pragma unroll 1
                    for (int sidx = idx_max; sidx > idx_max - n_chk && no_uturn; --sidx) {
                        const V cp = mg_ld<D, NS>(slot(S_CKP + sidx));
                        // span = rho_s - checkpoint rho + checkpoint momentum
                        V span = rho_s;
                        mg_axpy<D, NS>(span, T(-1), mg_ld<D, NS>(slot(S_CKR + sidx)));
                        mg_axpy<D, NS>(span, T(1), cp);
                        float D1, D2;
                        warp_dots2(mg_dot<D, NS>(span, cp, lane), mg_dot<D, NS>(span, p, lane), lane, D1, D2);
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
                    mg_st<D, NS>(slot(S_SQ), q);
                    mg_st<D, NS>(slot(S_SG), gq);
                    s_lp = lqv; s_en = energy;
                }
                w_s = w_new;
                accept_sum += nuts_exp_neg<T>(delta);
                if (divergent) not_div = false;
                alive = (!divergent) && no_uturn;
            }
            if (boot) {
                mg_st<D, NS>(slot(S_CQ), q);
                mg_st<D, NS>(slot(S_CG), gq);
                c_lp = lqv;
                break;
            }
            const double tree_w = alive ? w_s : -CUDART_INF;
            double thresh = tree_w - wsum;
            if (isnan(thresh)) thresh = 0.0;
            const double u = (double)rng_uniform<T>(key, (uint32_t)depth, (uint32_t)t, STREAM_NUTS_TOP);
            if ((nuts_log1m<T>(u) <= thresh) && alive) {
                mg_st<D, NS>(slot(S_CQ), mg_ld<D, NS>(slot(S_SQ)));
                mg_st<D, NS>(slot(S_CG), mg_ld<D, NS>(slot(S_SG)));
                c_lp = s_lp; c_energy = s_en; is_acc = true;
            }
            wsum = nuts_logaddexp<T>(wsum, tree_w);
            mg_st<D, NS>(slot(e_q), q);
            mg_st<D, NS>(slot(e_p), p);
            mg_st<D, NS>(slot(e_g), gq);
            V rho = mg_ld<D, NS>(slot(S_RHO));
            mg_axpy<D, NS>(rho, T(1), rho_s);
            mg_st<D, NS>(slot(S_RHO), rho);
            float D1, D2;
            warp_dots2(mg_dot<D, NS>(rho, mg_ld<D, NS>(slot(S_LP)), lane), mg_dot<D, NS>(rho, mg_ld<D, NS>(slot(S_RP)), lane), lane, D1, D2);
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
            mp.scatter(draws + (size_t)r * P, mg_ld<D, NS>(slot(S_CQ)), lane);
            if (lane == 0)
                write_stats<T>(stats + (size_t)r * SCCODA_NSTAT, lp_cur, lar, is_acc, eps, !not_div, leapfrogs, c_energy, cont);
        }
    }
    __syncwarp();
    mp.scatter(vec, mg_ld<D, NS>(slot(S_CQ)), lane);
    __syncwarp();
    store_chain_state<T>(a, chain, vec, ad, lp_cur, lane, 32);
}

}  // namespace sccoda
