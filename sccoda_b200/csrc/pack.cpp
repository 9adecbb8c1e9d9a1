// Host-side packer of libsccoda_b200.so: raw (x, y, reference) -> the packed dataset layout of include/sccoda_b200.h.
//
// It replaces the input handling of CompositionalModel.__init__ (sccoda/model/scCODA_model.py:90-100: cast to float64,
// zeros -> 0.5 pseudocount BEFORE the row sums) for B datasets at once and builds everything the kernels stage:
// rows sorted by covariate group, the element list, the pair / item layout of the gradient pass (grouped.cuh) and the
// parameter-independent constant of the log-density.  sccoda_b200/_pack.py: pack_numpy is the same layout in numpy
// (tests/test_pack_native.py compares the two); this one is what the product calls, threaded over datasets.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <new>
#include <numeric>
#include <thread>
#include <vector>

#include "../../include/sccoda_b200.h"

namespace sccoda_capi {
int set_error(int code, const char* fmt, ...);   // capi.cu: fills sccoda_last_error()
}

struct sccoda_packed {
    int B = 0, N = 0, K = 0, D = 0, Umax = 0, dtype = 0, grouped = 0;
    int NImax = 0, NOmax = 0, NGmax = 0, NF = 0, NT = 0, NOT = 0, nbig_min = 0;
    // float64 masters (the typed copies below are what sccoda_problem points at)
    std::vector<double> y, ntot, xu, ey, eycst, lconst, logp_const, iy, pcoef, oy;
    std::vector<float> ntot32, xu32, ey32, eycst32, iy32, pcoef32, oy32;
    std::vector<int32_t> grp, rstart, U, ref, colperm, nbig, row_order, NI, NG;
    std::vector<uint32_t> eidx, odesc;
    std::vector<uint16_t> ipair, ipos, opair, opos;
    std::vector<uint8_t> inreal;
};

namespace {

constexpr int R = SCCODA_ITEM_R;
constexpr double LOG_2PI = 1.8378770664093454835606594728112;

// numpy's float64 add.reduce over a contiguous axis: pairwise sum (8 accumulators, blocks of 128) — reproduced so that
// non-integer data give bit-identical row totals in both packers
double np_pairwise(const double* a, long n) {
    if (n < 8) {
        double r = 0.0;
        for (long i = 0; i < n; ++i) r += a[i];
        return r;
    }
    if (n <= 128) {
        double r[8];
        for (int j = 0; j < 8; ++j) r[j] = a[j];
        long i = 8;
        for (; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += a[i + j];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    }
    long n2 = n / 2;
    n2 -= n2 % 8;
    return np_pairwise(a, n2) + np_pairwise(a + n2, n - n2);
}
double np_sum(const double* a, long n) { return np_pairwise(a, n); }

// lgamma without the write to the global `signgam` (every thread of the packer calls it ~800 times per dataset: the
// shared cache line would bounce between the cores)
inline double lgamma_mt(double v) {
    int sign;
    return ::lgamma_r(v, &sign);
}

// An exception inside a worker (std::bad_alloc) must not escape its thread: it is recorded and rethrown by the caller's
// thread after the join (sccoda_pack turns it into an error code).
template <typename F>
void parallel_for(int n, int threads, F&& f) {
    if (threads <= 1 || n <= 1) {
        for (int i = 0; i < n; ++i) f(i);
        return;
    }
    if (threads > n) threads = n;
    std::vector<std::thread> pool;
    pool.reserve(threads);
    std::atomic<bool> failed{false};
    for (int t = 0; t < threads; ++t)
        pool.emplace_back([&, t]() {
            try {
                const int lo = (int)((long long)n * t / threads), hi = (int)((long long)n * (t + 1) / threads);
                for (int i = lo; i < hi && !failed.load(std::memory_order_relaxed); ++i) f(i);
            } catch (...) {
                failed.store(true);
            }
        });
    for (auto& th : pool) th.join();
    if (failed.load()) throw std::bad_alloc();
}

// items and odd counts of one dataset before the batch-wide padding (build_items of _pack.py)
struct DsItems {
    std::vector<double> iy;                 // item-major: R counts per item
    std::vector<int> iu, icol, itt, ireal;  // group, column (K = row totals), slot within the pair, data counts
    std::vector<double> oy;
    std::vector<int> ou, ocol, oslot, ofirst, ocount;
    int nslot = 1, nt = 1, not_ = 0;
};

struct DsScratch {
    std::vector<int> order, grp;            // packed row -> original row; group of each packed row
    int U = 0;
    DsItems it;
};

}  // namespace

extern "C" {

void sccoda_packed_free(sccoda_packed* pk) { delete pk; }

static int pack_impl(int32_t B, int32_t N, int32_t K, int32_t D, const double* x, const double* y, const int32_t* ref,
                     int32_t dtype, int32_t pseudocount, int32_t grouped_req, int32_t threads, sccoda_packed** out);

int sccoda_pack(int32_t B, int32_t N, int32_t K, int32_t D, const double* x, const double* y, const int32_t* ref,
                int32_t dtype, int32_t pseudocount, int32_t grouped_req, int32_t threads, sccoda_packed** out) {
    try {
        return pack_impl(B, N, K, D, x, y, ref, dtype, pseudocount, grouped_req, threads, out);
    } catch (const std::exception& e) {
        if (out) *out = nullptr;
        return sccoda_capi::set_error(-54, "sccoda_pack: %s", e.what());
    }
}

static int pack_impl(int32_t B, int32_t N, int32_t K, int32_t D, const double* x, const double* y, const int32_t* ref,
                     int32_t dtype, int32_t pseudocount, int32_t grouped_req, int32_t threads, sccoda_packed** out) {
    using sccoda_capi::set_error;
    if (!out) return set_error(-50, "sccoda_pack: out is NULL");
    *out = nullptr;
    if (!x || !y || !ref) return set_error(-50, "sccoda_pack: x / y / ref must not be NULL");
    if (B < 1 || N < 1 || K < 2 || D < 1) return set_error(-51, "sccoda_pack: need B >= 1, N >= 1, K >= 2, D >= 1 (got B=%d N=%d K=%d D=%d)", B, N, K, D);
    if (dtype != SCCODA_F32 && dtype != SCCODA_F64) return set_error(-5, "dtype must be SCCODA_F32 or SCCODA_F64");
    for (int b = 0; b < B; ++b)
        if (ref[b] < 0 || ref[b] >= K) return set_error(-52, "Reference index is not a valid cell type name or numerical index!");
    if (N > 65535) return set_error(-53, "packed element index needs Umax*K <= 65535 and N <= 65535");
    if (threads <= 0) {
        threads = (int)std::thread::hardware_concurrency();
        if (threads < 1) threads = 1;
        if (threads > 32) threads = 32;
    }
    std::unique_ptr<sccoda_packed> holder(new sccoda_packed());   // freed on every early return and on exceptions
    sccoda_packed* pk = holder.get();
    pk->B = B; pk->N = N; pk->K = K; pk->D = D; pk->dtype = dtype;
    const size_t NK = (size_t)N * K;
    std::vector<DsScratch> ds(B);
    pk->y.assign((size_t)B * NK, 0.0);
    pk->ntot.assign((size_t)B * N, 0.0);
    pk->grp.assign((size_t)B * N, 0);
    pk->row_order.assign((size_t)B * N, 0);
    pk->U.assign(B, 0);
    pk->ref.assign(ref, ref + B);
    pk->lconst.assign(B, 0.0);
    pk->logp_const.assign(B, 0.0);

    // ---- phase 1: covariate groups, row order, counts after the pseudocount, row totals
    parallel_for(B, threads, [&](int b) {
        const double* xb = x + (size_t)b * N * D;
        DsScratch& s = ds[b];
        std::vector<int> idx(N);
        std::iota(idx.begin(), idx.end(), 0);
        auto less = [&](int i, int j) {          // lexicographic order of the covariate rows (np.unique(axis=0))
            for (int d = 0; d < D; ++d) {
                const double a = xb[(size_t)i * D + d], c = xb[(size_t)j * D + d];
                if (a < c) return true;
                if (a > c) return false;
            }
            return false;
        };
        std::stable_sort(idx.begin(), idx.end(), less);   // equal rows keep their original order: stable argsort of the group ids
        s.order = idx;
        s.grp.resize(N);
        int g = 0;
        for (int n = 0; n < N; ++n) {
            if (n > 0 && (less(idx[n - 1], idx[n]) || less(idx[n], idx[n - 1]))) ++g;
            s.grp[n] = g;
        }
        s.U = g + 1;
        double* yb = pk->y.data() + (size_t)b * NK;
        for (int n = 0; n < N; ++n) {
            const double* src = y + ((size_t)b * N + idx[n]) * K;
            for (int k = 0; k < K; ++k) {
                double v = src[k];
                if (pseudocount && v == 0.0) v = 0.5;                        // scCODA_model.py:94-96
                yb[(size_t)n * K + k] = v;
            }
            pk->ntot[(size_t)b * N + n] = np_sum(yb + (size_t)n * K, K);     // :99-100, after the pseudocount
            pk->grp[(size_t)b * N + n] = s.grp[n];
            pk->row_order[(size_t)b * N + n] = idx[n];
        }
        pk->U[b] = s.U;
    });
    int Umax = 0;
    for (int b = 0; b < B; ++b) Umax = std::max(Umax, ds[b].U);
    pk->Umax = Umax;
    if ((long long)Umax * K > 65535) {
        return set_error(-53, "packed element index needs Umax*K <= 65535 and N <= 65535");
    }
    const int NP = Umax * (K + 1);
    pk->xu.assign((size_t)B * Umax * D, 0.0);
    pk->rstart.assign((size_t)B * (Umax + 1), N);
    pk->ey.assign((size_t)B * NK, 0.0);
    pk->eycst.assign((size_t)B * NK, 0.0);
    pk->eidx.assign((size_t)B * NK, 0u);
    pk->colperm.assign((size_t)B * K, 0);
    pk->nbig.assign(B, 0);
    pk->pcoef.assign((size_t)B * NP * 6, 0.0);
    const int DK1 = D * (K - 1);
    const double prior_const = D * std::log(2.0 / M_PI) - DK1 * LOG_2PI - K * (std::log(5.0) + 0.5 * LOG_2PI);

    // ---- phase 2: constants, element list (columns by decreasing minimum count), items of every pair
    parallel_for(B, threads, [&](int b) {
        DsScratch& s = ds[b];
        const double* yb = pk->y.data() + (size_t)b * NK;
        const double* nt = pk->ntot.data() + (size_t)b * N;
        const double* xb = x + (size_t)b * N * D;
        int32_t* rs = pk->rstart.data() + (size_t)b * (Umax + 1);
        for (int n = N - 1; n >= 0; --n) rs[s.grp[n]] = n;
        for (int n = 0; n < N; ++n)
            if (n == rs[s.grp[n]])
                for (int d = 0; d < D; ++d) pk->xu[((size_t)b * Umax + s.grp[n]) * D + d] = xb[(size_t)s.order[n] * D + d];
        // constants (fp64): multinomial coefficient, priors and — fp32 build — sum lgamma(y), see special.cuh
        double lc = 0.0, lgsum = 0.0;
        for (int n = 0; n < N; ++n) lc += lgamma_mt(nt[n] + 1.0);
        std::vector<double> ycst(NK, 0.0);
        double ly1 = 0.0;
        for (size_t e = 0; e < NK; ++e) {
            const double v = yb[e];
            ly1 += lgamma_mt(v + 1.0);
            if (dtype == SCCODA_F32) {
                const double lg = lgamma_mt(v);
                lgsum += lg;
                ycst[e] = v >= 6.0 ? lg - ((v - 0.5) * std::log(v) - v + 0.5 * LOG_2PI) : lg;   // r(y) / lgamma(y)
            }
        }
        pk->logp_const[b] = lc - ly1;
        pk->lconst[b] = prior_const + pk->logp_const[b] + (dtype == SCCODA_F32 ? lgsum : 0.0);
        // element list
        std::vector<double> colmin(K);
        for (int k = 0; k < K; ++k) {
            double m = yb[k];
            for (int n = 1; n < N; ++n) m = std::min(m, yb[(size_t)n * K + k]);
            colmin[k] = m;
        }
        std::vector<int> perm(K);
        std::iota(perm.begin(), perm.end(), 0);
        std::stable_sort(perm.begin(), perm.end(), [&](int a, int c) { return colmin[a] > colmin[c]; });
        int nbigcols = 0;
        for (int k = 0; k < K; ++k) nbigcols += colmin[k] >= 6.0;
        pk->nbig[b] = N * nbigcols;
        for (int kk = 0; kk < K; ++kk) {
            const int k = perm[kk];
            pk->colperm[(size_t)b * K + kk] = k;
            for (int n = 0; n < N; ++n) {
                const size_t e = (size_t)b * NK + (size_t)kk * N + n;
                pk->ey[e] = yb[(size_t)n * K + k];
                pk->eycst[e] = ycst[(size_t)n * K + k];
                pk->eidx[e] = (uint32_t)(s.grp[n] * K + k) | ((uint32_t)n << 16);
            }
        }
        // pair / item layout (build_items of _pack.py): per group, per column (K = the row totals)
        DsItems& it = s.it;
        std::vector<double> col;
        for (int u = 0; u < s.U; ++u) {
            const int r0 = rs[u], r1 = (u + 1 < s.U) ? rs[u + 1] : N;
            for (int c = 0; c <= K; ++c) {
                col.clear();
                double W[6] = {0, 0, 0, 0, 0, 0};
                int nb = 0, no = 0;
                const size_t odd_first = it.oy.size();
                for (int n = r0; n < r1; ++n) {
                    const double v = c < K ? yb[(size_t)n * K + c] : nt[n];
                    if (v >= 6.0) {
                        col.push_back(v);
                        ++nb;
                    } else if (v == std::floor(v) && v >= 1.0) {
                        for (int i = 0; i < 6; ++i) W[i] += v > i ? 1.0 : 0.0;
                    } else if (v != 0.0) {
                        it.oy.push_back(v);
                        ++no;
                    }
                }
                for (int i = 0; i < 6; ++i) W[i] += nb + no;
                const int pair = c < K ? u * K + c : Umax * K + u;
                double* pc = pk->pcoef.data() + ((size_t)b * NP + pair) * 6;
                pc[0] = W[0] + W[5]; pc[1] = 5.0 * W[0]; pc[2] = W[1] + W[4]; pc[3] = 4.0 * W[1] + W[4];
                pc[4] = W[2] + W[3]; pc[5] = 3.0 * W[2] + 2.0 * W[3];
                const int nf = (nb + R - 1) / R;
                it.nslot = std::max(it.nslot, nf + (no > 0 ? 1 : 0));
                it.nt = std::max(it.nt, nf);
                it.not_ = std::max(it.not_, no);
                std::sort(col.begin(), col.end(), [](double a, double c2) { return a > c2; });   // decreasing counts
                for (int tt = 0; tt < nf; ++tt) {
                    for (int j = 0; j < R; ++j) {
                        const int q = tt * R + j;
                        it.iy.push_back(q < nb ? col[q] : 6.0);
                    }
                    it.iu.push_back(u); it.icol.push_back(c); it.itt.push_back(tt);
                    it.ireal.push_back(std::min(R, nb - R * tt));
                }
                if (no > 0) {
                    it.ou.push_back(u); it.ocol.push_back(c); it.oslot.push_back(nf);
                    it.ofirst.push_back((int)odd_first); it.ocount.push_back(no);
                }
            }
        }
    });

    // ---- phase 3: batch-wide sizes, padded item arrays
    int nslot_max = 1, nt_max = 1, not_max = 0, ni_max = 0, no_max = 0, ng_max = 0;
    for (int b = 0; b < B; ++b) {
        const DsItems& it = ds[b].it;
        nslot_max = std::max(nslot_max, it.nslot); nt_max = std::max(nt_max, it.nt); not_max = std::max(not_max, it.not_);
        ni_max = std::max(ni_max, (int)it.itt.size()); no_max = std::max(no_max, (int)it.oy.size());
        ng_max = std::max(ng_max, (int)it.ocount.size());
    }
    int NF = 2;
    while (NF < nslot_max) NF *= 2;
    int NImax = std::max(2, ni_max + (ni_max & 1)), NOmax = std::max(1, no_max), NGmax = std::max(1, ng_max);
    bool items_ok = !((long long)NP * NF > 65535 || NImax > 65535 || NOmax > 65535);
    if (!items_ok && grouped_req > 0) {
        return set_error(-55, "pair/item layout needs Umax*(K+1)*NF <= 65535 and at most 65535 items / odd counts");
    }
    pk->NI.assign(B, 0);
    pk->NG.assign(B, 0);
    if (!items_ok) {
        // too many pairs / items for the 16-bit indices of the pair/item layout: element layout, placeholder item arrays
        NF = 2; NImax = 2; NOmax = 1; NGmax = 1; nt_max = 1; not_max = 0;
        std::fill(pk->pcoef.begin(), pk->pcoef.end(), 0.0);
        pk->iy.assign((size_t)B * R * NImax, 6.0);
        pk->ipair.assign((size_t)B * NImax, 0); pk->ipos.assign((size_t)B * NImax, 0); pk->inreal.assign((size_t)B * NImax, 0);
        pk->oy.assign((size_t)B * NOmax, 1.0);
        pk->opair.assign((size_t)B * NGmax, 0); pk->opos.assign((size_t)B * NGmax, 0); pk->odesc.assign((size_t)B * NGmax, 0u);
    } else {
        pk->iy.assign((size_t)B * R * NImax, 6.0);
        pk->ipair.assign((size_t)B * NImax, 0);
        pk->ipos.assign((size_t)B * NImax, (uint16_t)(NP * NF));        // padding items write to a spare slot
        pk->inreal.assign((size_t)B * NImax, 0);
        pk->oy.assign((size_t)B * NOmax, 1.0);
        pk->opair.assign((size_t)B * NGmax, 0);
        pk->opos.assign((size_t)B * NGmax, (uint16_t)(NP * NF));
        pk->odesc.assign((size_t)B * NGmax, 0u);
        parallel_for(B, threads, [&](int b) {
            const DsItems& it = ds[b].it;
            const int ni = (int)it.itt.size(), ng = (int)it.ocount.size();
            pk->NI[b] = ni;
            pk->NG[b] = ng;
            for (int i = 0; i < ni; ++i) {
                const int pair = it.icol[i] < K ? it.iu[i] * K + it.icol[i] : Umax * K + it.iu[i];
                for (int j = 0; j < R; ++j) pk->iy[((size_t)b * R + j) * NImax + i] = it.iy[(size_t)i * R + j];
                pk->ipair[(size_t)b * NImax + i] = (uint16_t)pair;
                pk->ipos[(size_t)b * NImax + i] = (uint16_t)(pair * NF + it.itt[i]);
                pk->inreal[(size_t)b * NImax + i] = (uint8_t)it.ireal[i];
            }
            std::copy(it.oy.begin(), it.oy.end(), pk->oy.begin() + (size_t)b * NOmax);
            for (int q = 0; q < ng; ++q) {
                const int pair = it.ocol[q] < K ? it.ou[q] * K + it.ocol[q] : Umax * K + it.ou[q];
                pk->opair[(size_t)b * NGmax + q] = (uint16_t)pair;
                pk->opos[(size_t)b * NGmax + q] = (uint16_t)(pair * NF + it.oslot[q]);
                pk->odesc[(size_t)b * NGmax + q] = (uint32_t)it.ofirst[q] | ((uint32_t)it.ocount[q] << 16);
            }
        });
    }
    pk->NF = NF; pk->NImax = NImax; pk->NOmax = NOmax; pk->NGmax = NGmax; pk->NT = nt_max; pk->NOT = not_max;

    // ---- layout choice (see _pack.py: pack): pair/item gradient for the designs the lane-per-cell-type kernels take and
    // whenever the item padding stays below 60 % of the counts
    int grouped = grouped_req > 0 ? 1 : 0;
    if (grouped_req < 0) {
        const bool lane_owned = Umax <= 32 && ((K <= 31 && D <= 4) || (K <= 62 && D <= 3));
        int ni = 0;
        for (int b = 0; b < B; ++b) ni = std::max(ni, pk->NI[b]);
        grouped = items_ok && (lane_owned || (double)R * ni <= 1.6 * N * (K + 1)) ? 1 : 0;
    }
    if (!items_ok) grouped = 0;

    auto typed_copies = [&]() {
        if (dtype != SCCODA_F32) return;
        struct Job { const std::vector<double>* s; std::vector<float>* d; };
        Job jobs[] = {{&pk->ntot, &pk->ntot32}, {&pk->xu, &pk->xu32}, {&pk->ey, &pk->ey32}, {&pk->eycst, &pk->eycst32},
                      {&pk->iy, &pk->iy32}, {&pk->pcoef, &pk->pcoef32}, {&pk->oy, &pk->oy32}};
        for (auto& j : jobs) j.d->resize(j.s->size());
        const int chunks = threads * 4;                       // every array in `chunks` slices, converted in parallel
        parallel_for(chunks, threads, [&](int ch) {
            for (auto& j : jobs) {
                const size_t n = j.s->size(), lo = n * ch / chunks, hi = n * (ch + 1) / chunks;
                const double* s = j.s->data();
                float* d = j.d->data();
                for (size_t i = lo; i < hi; ++i) d[i] = (float)s[i];
            }
        });
    };
    auto set_nbig_min = [&]() {
        int m = pk->nbig[0];
        for (int b = 1; b < B; ++b) m = std::min(m, pk->nbig[b]);
        pk->nbig_min = m;
    };

    if (grouped && grouped_req < 0) {
        // The automatic choice asks the library's own launch planner (NUTS with the default tree depth = the largest
        // footprint) whether the pair/item arrays of a dataset fit one CTA's shared memory; very long datasets take the
        // element layout.  The plan depends on the sizes only (and on nbig_min of the grouped list order = the smallest
        // number of counts >= 6 of a dataset), not on the order of the element list: asked before anything is reordered.
        int nb_min = (int)NK;
        for (int b = 0; b < B; ++b) {
            const double* ey = pk->ey.data() + (size_t)b * NK;
            int nb = 0;
            for (size_t e = 0; e < NK; ++e) nb += ey[e] >= 6.0;
            nb_min = std::min(nb_min, nb);
        }
        sccoda_problem prob;
        pk->grouped = 1;
        pk->nbig_min = nb_min;
        sccoda_packed_problem(pk, 1, &prob);
        const int dummy = 0;                                  // the planner only checks that the arrays are there
        const void* any = &dummy;
        prob.ey = prob.eycst = prob.ntot = prob.xu = prob.iy = prob.pcoef = prob.oy = any;
        sccoda_sampler smp;
        std::memset(&smp, 0, sizeof(smp));
        smp.method = SCCODA_METHOD_NUTS; smp.num_results = 1; smp.num_leapfrog = 10; smp.max_tree_depth = 10;
        smp.step_size = 0.01; smp.target_accept = 0.75; smp.store_draws = 1;
        int32_t w, c, g, sm;
        if (sccoda_launch_plan(&prob, &smp, &w, &c, &g, &sm) == -9) grouped = 0;
    }
    if (grouped) {
        // only the value pass walks the element list of the grouped layout: every count >= 6 first (stable partition)
        parallel_for(B, threads, [&](int b) {
            double* ey = pk->ey.data() + (size_t)b * NK;
            double* ec = pk->eycst.data() + (size_t)b * NK;
            uint32_t* ei = pk->eidx.data() + (size_t)b * NK;
            std::vector<double> ty(ey, ey + NK), tc(ec, ec + NK);
            std::vector<uint32_t> ti(ei, ei + NK);
            size_t o = 0;
            for (int pass = 0; pass < 2; ++pass)
                for (size_t e = 0; e < NK; ++e)
                    if ((ty[e] >= 6.0) == (pass == 0)) { ey[o] = ty[e]; ec[o] = tc[e]; ei[o] = ti[e]; ++o; }
            int nb = 0;
            for (size_t e = 0; e < NK; ++e) nb += ey[e] >= 6.0;
            pk->nbig[b] = nb;
        });
        pk->grouped = 1;
        set_nbig_min();
        typed_copies();
    } else {
        pk->grouped = 0;
        set_nbig_min();
        typed_copies();
    }
    *out = holder.release();
    return 0;
}

int sccoda_packed_problem(const sccoda_packed* pk, int32_t C, sccoda_problem* p) {
    if (!pk || !p) return sccoda_capi::set_error(-50, "sccoda_packed_problem: NULL argument");
    std::memset(p, 0, sizeof(*p));
    p->B = pk->B; p->C = C; p->N = pk->N; p->K = pk->K; p->D = pk->D; p->Umax = pk->Umax; p->dtype = pk->dtype;
    p->nbig_min = pk->nbig_min; p->grouped = pk->grouped; p->NImax = pk->NImax; p->NOmax = pk->NOmax; p->NGmax = pk->NGmax;
    p->NF = pk->NF; p->NT = pk->NT; p->NOT = pk->NOT;
    const bool f32 = pk->dtype == SCCODA_F32;
    p->ey = f32 ? (const void*)pk->ey32.data() : (const void*)pk->ey.data();
    p->eycst = f32 ? (const void*)pk->eycst32.data() : (const void*)pk->eycst.data();
    p->ntot = f32 ? (const void*)pk->ntot32.data() : (const void*)pk->ntot.data();
    p->xu = f32 ? (const void*)pk->xu32.data() : (const void*)pk->xu.data();
    p->iy = f32 ? (const void*)pk->iy32.data() : (const void*)pk->iy.data();
    p->pcoef = f32 ? (const void*)pk->pcoef32.data() : (const void*)pk->pcoef.data();
    p->oy = f32 ? (const void*)pk->oy32.data() : (const void*)pk->oy.data();
    p->eidx = pk->eidx.data(); p->colperm = pk->colperm.data(); p->nbig = pk->nbig.data(); p->grp = pk->grp.data();
    p->rstart = pk->rstart.data(); p->U = pk->U.data(); p->ref = pk->ref.data(); p->lconst = pk->lconst.data();
    p->NI = pk->NI.data(); p->ipair = pk->ipair.data(); p->ipos = pk->ipos.data(); p->inreal = pk->inreal.data();
    p->NG = pk->NG.data(); p->opair = pk->opair.data(); p->opos = pk->opos.data(); p->odesc = pk->odesc.data();
    return 0;
}

const int32_t* sccoda_packed_row_order(const sccoda_packed* pk) { return pk ? pk->row_order.data() : nullptr; }
const double* sccoda_packed_y(const sccoda_packed* pk) { return pk ? pk->y.data() : nullptr; }
const double* sccoda_packed_logp_const(const sccoda_packed* pk) { return pk ? pk->logp_const.data() : nullptr; }

}  // extern "C"
