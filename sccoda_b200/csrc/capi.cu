// C ABI (extern "C") of libsccoda_b200.so — see include/sccoda_b200.h.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <chrono>
#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/sccoda_b200.h"
#include "kernels.h"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CUDA_TRY(expr)                                                                       \
    do {                                                                                     \
        cudaError_t e__ = (expr);                                                            \
        if (e__ != cudaSuccess) return fail(-100, "%s: %s", #expr, cudaGetErrorString(e__)); \
    } while (0)

}  // namespace

namespace sccoda_capi {
// error text for the translation units outside this file (pack.cpp)
int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
}  // namespace sccoda_capi

namespace {

uint32_t magic_for(int d) { return d <= 1 ? 0u : (uint32_t)((0x100000000ull + (uint64_t)d - 1) / (uint64_t)d); }

int pow2_ceil(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}
int pow2_floor(int v) {
    int p = 1;
    while (p * 2 <= v) p <<= 1;
    return p;
}

int validate(const sccoda_problem* p) {
    if (!p) return fail(-1, "problem is NULL");
    if (p->B < 1 || p->C < 1) return fail(-2, "need B >= 1 and C >= 1 (got B=%d C=%d)", p->B, p->C);
    if (p->N < 1 || p->K < 2 || p->D < 1) return fail(-3, "need N >= 1, K >= 2, D >= 1 (got N=%d K=%d D=%d)", p->N, p->K, p->D);
    if (p->D > 32) return fail(-3, "need D <= 32 covariates (got D=%d)", p->D);
    if (p->Umax < 1 || p->Umax > p->N) return fail(-4, "need 1 <= Umax <= N (got Umax=%d N=%d)", p->Umax, p->N);
    if (p->dtype != SCCODA_F32 && p->dtype != SCCODA_F64) return fail(-5, "dtype must be SCCODA_F32 or SCCODA_F64");
    if ((uint64_t)p->N * p->K * p->K >= 0x100000000ull) return fail(-6, "N*K*K must be < 2^32");
    if ((uint64_t)p->Umax * p->K > 65535ull || p->N > 65535) return fail(-6, "need Umax*K <= 65535 and N <= 65535 (packed element index)");
    if (p->nbig_min < 0 || p->nbig_min > p->N * p->K) return fail(-6, "nbig_min out of range");
    if (!p->ey || !p->eycst || !p->eidx || !p->colperm || !p->nbig || !p->ntot || !p->xu || !p->grp || !p->rstart ||
        !p->U || !p->ref || !p->lconst)
        return fail(-7, "problem has a NULL array");
    if (p->grouped) {
        if (p->NImax < 2 || (p->NImax & 1) || p->NImax > 65535 || p->NOmax < 1 || p->NOmax > 65535 || p->NGmax < 1)
            return fail(-6, "grouped layout needs an even 2 <= NImax <= 65535, 1 <= NOmax <= 65535 and NGmax >= 1 (got %d, %d, %d)",
                        p->NImax, p->NOmax, p->NGmax);
        if (p->NT < 1 || p->NT > p->NF || p->NOT < 0) return fail(-6, "grouped layout needs 1 <= NT <= NF and NOT >= 0 (got %d, %d)", p->NT, p->NOT);
        if (p->NF < 2 || (p->NF & (p->NF - 1))) return fail(-6, "grouped layout needs NF = power of two >= 2 (got %d)", p->NF);
        if ((uint64_t)p->Umax * (p->K + 1) * p->NF > 65535ull) return fail(-6, "grouped layout needs Umax*(K+1)*NF <= 65535");
        if (!p->NI || !p->iy || !p->ipair || !p->ipos || !p->inreal || !p->pcoef || !p->NG || !p->oy || !p->opair || !p->opos || !p->odesc)
            return fail(-7, "grouped problem has a NULL pair/item array");
    }
    return 0;
}

// sccoda_sampler::reject_conc_above: 0 = the default of the arithmetic type, negative or infinite = no limit
double default_conc_max(int dtype) { return dtype == SCCODA_F32 ? 1e6 : 1e15; }
double conc_max_from(double reject_conc_above, int dtype) {
    if (reject_conc_above == 0.0) return default_conc_max(dtype);
    if (reject_conc_above < 0.0) return (double)INFINITY;
    return reject_conc_above;
}

void fill_dims(const sccoda_problem* p, sccoda::KArgs& a) {
    a.dm.N = p->N; a.dm.K = p->K; a.dm.D = p->D;
    a.dm.K1 = p->K - 1; a.dm.DK1 = p->D * (p->K - 1); a.dm.NK = p->N * p->K;
    a.dm.P = p->D + 2 * a.dm.DK1 + p->K;
    a.dm.magicK = magic_for(p->K);
    a.dm.magicK1 = magic_for(p->K - 1);  // 0 encodes a divisor of 1 (K == 2)
    a.dm.conc_max = default_conc_max(p->dtype);
    a.B = p->B; a.C = p->C; a.Umax = p->Umax;
    a.ey = p->ey; a.eycst = p->eycst; a.eidx = p->eidx; a.colperm = p->colperm; a.nbig = p->nbig;
    a.ntot = p->ntot; a.xu = p->xu;
    a.grp = p->grp; a.rstart = p->rstart; a.U = p->U; a.ref = p->ref; a.lconst = p->lconst;
    a.grouped = p->grouped ? 1 : 0;
    // entries of the element list the value pass wants in shared memory: all of them, or (grouped layout: the counts
    // >= 6 are read from the items) only the tail of counts below 6
    a.el_count = a.grouped ? a.dm.NK - p->nbig_min : a.dm.NK;
    a.NImax = a.grouped ? p->NImax : 0; a.NOmax = a.grouped ? p->NOmax : 0;
    a.NGmax = a.grouped ? p->NGmax : 0; a.NF = a.grouped ? p->NF : 0;
    a.NT = a.grouped ? p->NT : 0; a.NOT = a.grouped ? p->NOT : 0;
    a.NI = p->NI; a.iy = p->iy; a.ipair = p->ipair; a.ipos = p->ipos; a.inreal = p->inreal; a.pcoef = p->pcoef;
    a.NG = p->NG; a.oy = p->oy; a.opair = p->opair; a.opos = p->opos; a.odesc = p->odesc;
}

struct Plan {
    int W, cpc, grid, smem;
};

int make_plan(const sccoda_problem* p, sccoda::KArgs& a, int& kind, int W_req, int cpc_req, Plan* out) {
    int dev = 0, max_smem = 0, n_sm = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
        cudaGetLastError();
        max_smem = 227 * 1024;  // B200; used only for planning without a device
        n_sm = 148;
    }
    const int NK = p->N * p->K;
    const int kind_in = kind;
    // Two attempts: the lane-per-cell-type kernels where the shape allows them, then the generic kernels (a dataset
    // whose staged tile does not fit one CTA's shared memory in the specialised layout still runs)
    for (int attempt = 0; attempt < 2; ++attempt) {
        kind = kind_in;
        if (attempt == 1) a.generic_only = 1;
        int W = W_req;
        if (W == 0 && sccoda::select_kind(a, kind, p->dtype, 1) != kind &&
            (p->B * p->C >= 148 * 8 || (p->NT <= 8 && p->Umax <= 16))) {
            // case/control and multi-group shapes: the lane-per-cell-type kernels (one warp per chain) unless a few
            // chains face very long groups (more than 40 rows per group: the items of a pair are walked by one lane)
            // or very many groups (the group pairs are walked by one warp): the generic kernels spread those over
            // several warps per chain, which wins while the device is not full
            W = 1;
        }
        if (W == 0) {
            const int n_chains = p->B * p->C;
            const int w_work = pow2_ceil((NK + 511) / 512);
            const int fill = (148 * 8) / n_chains;  // warps per chain that would give ~8 warps per SM
            const int w_fill = pow2_floor(fill > 0 ? fill : 1);
            const int w_cap = pow2_ceil((NK + 31) / 32);  // never fewer than one element per thread
            const int w_lim = w_fill < w_cap ? w_fill : w_cap;
            W = w_work > w_lim ? w_work : w_lim;
            if (W > 16) W = 16;
            if (W < 1) W = 1;
        }
        if (W != 1 && W != 2 && W != 4 && W != 8 && W != 16) return fail(-8, "warps_per_chain must be 0,1,2,4,8,16");
        kind = sccoda::select_kind(a, kind, p->dtype, W);   // specialised kernels (may need less shared memory)
        const bool specialised = kind != kind_in;
        int cpc = cpc_req > 0 ? cpc_req : 128 / (32 * W);
        const int cpc_cap = (W <= 4 ? 128 : 32 * W) / (32 * W);   // launch bounds of the kernels
        if (cpc > cpc_cap) cpc = cpc_cap;
        if (cpc < 1) cpc = 1;
        if (cpc > p->C) cpc = p->C;
        // element list in shared memory: fp32 build only, and only while one chain group still fits
        a.el_smem = (p->dtype == SCCODA_F32) ? 1 : 0;
        if (a.el_smem && (int)sccoda::smem_bytes_for(a, kind, W, 1, p->dtype) > max_smem) a.el_smem = 0;
        while (cpc > 1 && (int)sccoda::smem_bytes_for(a, kind, W, cpc, p->dtype) > max_smem) cpc -= 1;
        const size_t smem = sccoda::smem_bytes_for(a, kind, W, cpc, p->dtype);
        if ((int)smem > max_smem) {
            if (specialised && attempt == 0 && !a.generic_only) continue;
            return fail(-9, "problem needs %zu bytes of shared memory per chain group, device limit is %d", smem, max_smem);
        }
        out->W = W; out->cpc = cpc; out->smem = (int)smem;
        out->grid = p->B * ((p->C + cpc - 1) / cpc);
        // A launch with more than 8 chains per SM on the case/control HMC kernel: one wide CTA per SM holding a balanced
        // share of the chains, its warps meeting at a barrier every 16 transitions (hmc_cc.cuh: WIDE; up to 16 chains
        // per SM with a 128-register build, beyond that the 72-register build that keeps 28 warps resident)
        a.wide = 0;
        const long long n_chains = (long long)p->B * p->C;
        // warps one SM holds of this kernel (its register budget): 28 for the case/control kernel, 4 x the resident CTAs
        // of 128 threads the multi-group kernel is compiled for
        const int wmax = kind == sccoda::KIND_HMC_CC ? 28 : (kind == sccoda::KIND_HMC_MG ? 4 * sccoda::mg_hmc_ctas(p->D, p->K <= 31 ? 1 : 2) : 0);
        // chains per SM from which the wide launch pays (measured: 10 and 12 chains per SM +3 %, 14 ... 16 +9 %, 21 ... 28
        // +15 ... 18 %; 7 and 8 chains per SM -2 %)
        const long long wide_from = kind == sccoda::KIND_HMC_CC ? 8 : (wmax * 2) / 3;
        // ... and only when the CTAs of the launch fill whole waves: one wave, or several with the last one >= 90 % full
        // (a CTA holds the SM alone; a sparse last wave costs more than the fair scheduling gains)
        const long long per_wave = (long long)n_sm * (wmax > 0 ? wmax : 1);
        const long long last_wave = n_chains % per_wave;
        const bool whole_waves = n_chains <= per_wave || last_wave == 0 || last_wave * 10 >= per_wave * 9;
        if (wmax > 0 && cpc_req == 0 && p->C <= wmax && n_chains > (long long)n_sm * wide_from && whole_waves) {
            const int per_sm = (int)((n_chains + n_sm - 1) / n_sm);
            int warps = per_sm < wmax ? per_sm : wmax;                        // 72 registers x 28 warps fill the register file
            a.wide = warps;
            a.wide_grid = per_sm <= wmax ? n_sm : (int)((n_chains + warps - 1) / warps);   // one CTA per SM, or several waves
            while (warps > 12 && (int)sccoda::smem_bytes_for(a, kind, W, cpc, p->dtype) > max_smem) {
                a.wide = --warps;                                             // the staged datasets of a CTA must fit (few chains per dataset)
                a.wide_grid = (int)((n_chains + warps - 1) / warps);
            }
            if ((int)sccoda::smem_bytes_for(a, kind, W, cpc, p->dtype) <= max_smem) {
                out->cpc = warps;
                out->grid = a.wide_grid;
                out->smem = (int)sccoda::smem_bytes_for(a, kind, W, cpc, p->dtype);
            } else {
                a.wide = 0;
            }
        }
        return 0;
    }
    return fail(-9, "problem does not fit the shared memory of one CTA");
}

int fill_sampler(const sccoda_sampler* s, sccoda::KArgs& a, int* kind, int dtype) {
    if (!s) return fail(-10, "sampler is NULL");
    if (s->method < 0 || s->method > 2) return fail(-11, "unknown method %d", s->method);
    if (s->num_results < 1 || s->num_burnin < 0) return fail(-12, "need num_results >= 1 and num_burnin >= 0");
    if (s->method != SCCODA_METHOD_NUTS && s->num_leapfrog < 1) return fail(-13, "need num_leapfrog >= 1");
    if (s->method == SCCODA_METHOD_NUTS && (s->max_tree_depth < 1 || s->max_tree_depth > 20))
        return fail(-14, "need 1 <= max_tree_depth <= 20");
    if (!(s->step_size > 0.0)) return fail(-15, "need step_size > 0");
    a.method = s->method; a.num_results = s->num_results; a.num_burnin = s->num_burnin; a.num_adapt = s->num_adapt;
    a.num_leapfrog = s->num_leapfrog; a.max_depth = s->method == SCCODA_METHOD_NUTS ? s->max_tree_depth : 0;
    a.step_begin = s->step_begin;
    a.step_end = s->step_end > 0 ? s->step_end : s->num_burnin + s->num_results;
    if (a.step_begin < 0 || a.step_end > s->num_burnin + s->num_results || a.step_begin > a.step_end)
        return fail(-16, "bad step range [%d,%d)", a.step_begin, a.step_end);
    a.store_draws = s->store_draws;
    a.generic_only = s->generic_only ? 1 : 0;
    a.dm.freeze = s->freeze_spike_slab ? 1 : 0;
    if (s->reject_conc_above != s->reject_conc_above) return fail(-17, "reject_conc_above is NaN");
    a.dm.conc_max = conc_max_from(s->reject_conc_above, dtype);
    a.step_size = s->step_size; a.target_accept = s->target_accept; a.seed = s->seed; a.chain_id0 = s->chain_id0;
    *kind = s->method == SCCODA_METHOD_NUTS ? sccoda::KIND_NUTS : sccoda::KIND_HMC;
    return 0;
}

// Device workspace of sccoda_sample_host, kept across calls (per device): the draw buffer of a sweep is tens of
// GB and cudaMalloc/cudaFree of it would otherwise sit inside every end-to-end call.
struct Workspace {
    void* ptr = nullptr;
    size_t bytes = 0;
    int device = -1;
};
Workspace g_ws;
std::mutex g_ws_mutex;   // sccoda_sample_host calls share the cached workspace: one at a time

int workspace_get(int device, size_t bytes, void** out) {
    if (g_ws.ptr != nullptr && (g_ws.device != device || g_ws.bytes < bytes)) {
        cudaFree(g_ws.ptr);
        g_ws = Workspace{};
    }
    if (g_ws.ptr == nullptr) {
        if (cudaMalloc(&g_ws.ptr, bytes) != cudaSuccess) {
            cudaGetLastError();
            g_ws = Workspace{};
            return fail(-42, "cudaMalloc(workspace, %zu bytes) failed", bytes);
        }
        g_ws.bytes = bytes;
        g_ws.device = device;
    }
    *out = g_ws.ptr;
    return 0;
}

size_t align256(size_t v) { return (v + 255) & ~size_t(255); }

}  // namespace

extern "C" {

int sccoda_version(void) { return 100; }

int sccoda_release_workspace(void) {
    std::lock_guard<std::mutex> lock(g_ws_mutex);
    if (g_ws.ptr != nullptr) cudaFree(g_ws.ptr);
    g_ws = Workspace{};
    return 0;
}

const char* sccoda_last_error(void) { return g_err; }

int sccoda_launch_plan(const sccoda_problem* prob, const sccoda_sampler* smp, int32_t* warps_per_chain,
                       int32_t* chains_per_cta, int32_t* grid, int32_t* smem_bytes) {
    int rc = validate(prob);
    if (rc) return rc;
    sccoda::KArgs a{};
    fill_dims(prob, a);
    int kind = 0;
    rc = fill_sampler(smp, a, &kind, prob->dtype);
    if (rc) return rc;
    Plan pl;
    rc = make_plan(prob, a, kind, smp->warps_per_chain, smp->chains_per_cta, &pl);
    if (rc) return rc;
    if (warps_per_chain) *warps_per_chain = pl.W;
    if (chains_per_cta) *chains_per_cta = pl.cpc;
    if (grid) *grid = pl.grid;
    if (smem_bytes) *smem_bytes = pl.smem;
    return 0;
}

int sccoda_kernel_kind(const sccoda_problem* prob, const sccoda_sampler* smp, int32_t* kind_out) {
    int rc = validate(prob);
    if (rc) return rc;
    sccoda::KArgs a{};
    fill_dims(prob, a);
    int kind = 0;
    rc = fill_sampler(smp, a, &kind, prob->dtype);
    if (rc) return rc;
    Plan pl;
    rc = make_plan(prob, a, kind, smp->warps_per_chain, smp->chains_per_cta, &pl);
    if (rc) return rc;
    if (kind_out) *kind_out = kind;
    return 0;
}

int sccoda_logp_grad(const sccoda_problem* prob, const void* theta, double* logp, void* grad, void* stream) {
    int rc = validate(prob);
    if (rc) return rc;
    if (!theta || !logp || !grad) return fail(-20, "theta/logp/grad must not be NULL");
    sccoda::KArgs a{};
    fill_dims(prob, a);
    a.init = theta; a.draws = grad; a.carry = logp;
    Plan pl;
    int kind = sccoda::KIND_LOGP;
    rc = make_plan(prob, a, kind, 0, 0, &pl);
    if (rc) return rc;
    CUDA_TRY(sccoda::launch_kernel(kind, prob->dtype, pl.W, pl.cpc, a, static_cast<cudaStream_t>(stream)));
    return 0;
}

int sccoda_logp_grad_as(const sccoda_problem* prob, int32_t kernel_kind, double reject_conc_above, const void* theta,
                        double* logp, void* grad, void* stream) {
    int rc = validate(prob);
    if (rc) return rc;
    if (!theta || !logp || !grad) return fail(-20, "theta/logp/grad must not be NULL");
    if (reject_conc_above != reject_conc_above) return fail(-17, "reject_conc_above is NaN");
    sccoda::KArgs a{};
    fill_dims(prob, a);
    a.dm.conc_max = conc_max_from(reject_conc_above, prob->dtype);
    a.init = theta; a.draws = grad; a.carry = logp;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (kernel_kind == sccoda::KIND_HMC || kernel_kind == sccoda::KIND_NUTS || kernel_kind == 0) {
        a.generic_only = 1;
        Plan pl;
        int kind = sccoda::KIND_LOGP;
        rc = make_plan(prob, a, kind, 0, 0, &pl);
        if (rc) return rc;
        CUDA_TRY(sccoda::launch_kernel(kind, prob->dtype, pl.W, pl.cpc, a, st));
        return 0;
    }
    const bool cc = kernel_kind == sccoda::KIND_HMC_CC || kernel_kind == sccoda::KIND_NUTS_CC;
    const bool mg = kernel_kind == sccoda::KIND_HMC_MG || kernel_kind == sccoda::KIND_NUTS_MG;
    if (!cc && !mg) return fail(-24, "unknown kernel kind %d", kernel_kind);
    if (cc ? !sccoda::fits_cc(a, prob->dtype) : !sccoda::fits_mg(a, prob->dtype))
        return fail(-25, "the problem does not have the shape kernel kind %d takes (fp32, pair/item layout, %s)", kernel_kind,
                    cc ? "D == 1, two covariate groups, K <= 31" : "<= 32 covariate groups, D <= 4 with K <= 31 or D <= 3 with K <= 62");
    int dev = 0, max_smem = 227 * 1024;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    const int kk = cc ? sccoda::KIND_HMC_CC : sccoda::KIND_HMC_MG;
    int cpc = prob->C < 4 ? prob->C : 4;
    while (cpc > 1 && (int)sccoda::smem_bytes_for(a, kk, 1, cpc, prob->dtype) > max_smem) cpc -= 1;
    const size_t smem = sccoda::smem_bytes_for(a, kk, 1, cpc, prob->dtype);
    if ((int)smem > max_smem) return fail(-9, "problem needs %zu bytes of shared memory per chain group, device limit is %d", smem, max_smem);
    const int grid = prob->B * ((prob->C + cpc - 1) / cpc);
    CUDA_TRY(cc ? sccoda::launch_logp_cc(a, grid, cpc * 32, smem, st) : sccoda::launch_logp_mg(a, grid, cpc * 32, smem, st));
    return 0;
}

int sccoda_sample(const sccoda_problem* prob, const sccoda_sampler* smp, const void* init, void* draws, void* stats,
                  double* carry, void* stream) {
    int rc = validate(prob);
    if (rc) return rc;
    sccoda::KArgs a{};
    fill_dims(prob, a);
    int kind = 0;
    rc = fill_sampler(smp, a, &kind, prob->dtype);
    if (rc) return rc;
    if (a.step_begin == 0 && !init) return fail(-21, "init must not be NULL when step_begin == 0");
    if (a.step_begin > 0 && !carry) return fail(-22, "carry must not be NULL when resuming (step_begin > 0)");
    if (a.store_draws && (!draws || !stats)) return fail(-23, "draws/stats must not be NULL when store_draws != 0");
    a.init = init; a.draws = draws; a.stats = stats; a.carry = carry;
    Plan pl;
    rc = make_plan(prob, a, kind, smp->warps_per_chain, smp->chains_per_cta, &pl);
    if (rc) return rc;
    CUDA_TRY(sccoda::launch_kernel(kind, prob->dtype, pl.W, pl.cpc, a, static_cast<cudaStream_t>(stream)));
    return 0;
}

int sccoda_summarize(const sccoda_problem* prob, int32_t num_results, int32_t skip, const void* draws,
                     const void* stats, double* summary, void* stream) {
    int rc = validate(prob);
    if (rc) return rc;
    if (!draws || !stats || !summary) return fail(-30, "draws/stats/summary must not be NULL");
    if (skip < 0 || skip > num_results) return fail(-31, "need 0 <= skip <= num_results");
    CUDA_TRY(sccoda::launch_summary(prob->dtype, prob->B, prob->C, prob->N, prob->K, prob->D, prob->ref, num_results, skip,
                                    draws, stats, summary, static_cast<cudaStream_t>(stream)));
    return 0;
}

int sccoda_predict(const sccoda_problem* prob, const int32_t* row_order, int32_t num_results, int32_t draw_begin,
                   int32_t draw_end, const void* draws, void* concentration, void* prediction, void* stream) {
    int rc = validate(prob);
    if (rc) return rc;
    if (!draws || (!concentration && !prediction)) return fail(-32, "draws and at least one output must not be NULL");
    if (draw_begin < 0 || draw_end > num_results || draw_begin >= draw_end)
        return fail(-33, "need 0 <= draw_begin < draw_end <= num_results");
    cudaError_t e = sccoda::launch_predict(prob->dtype, prob->B, prob->C, prob->N, prob->K, prob->D, prob->Umax, prob->ref, prob->xu,
                                           prob->grp, prob->ntot, row_order, num_results, draw_begin, draw_end, draws, concentration,
                                           prediction, static_cast<cudaStream_t>(stream));
    if (e == cudaErrorInvalidConfiguration)
        return fail(-9, "problem needs more shared memory per warp than the device has (Umax*K too large)");
    CUDA_TRY(e);
    return 0;
}

int sccoda_sample_host(const sccoda_problem* ph, const sccoda_sampler* smp, const void* init, void* draws, void* stats,
                       double* summary, int32_t summary_skip, int32_t device, float* kernel_ms) {
    if (!ph || !smp || !init) return fail(-40, "NULL argument");
    if (ph->B < 1 || ph->C < 1 || ph->N < 1 || ph->K < 2 || ph->D < 1 || ph->Umax < 1) return fail(-41, "bad problem dimensions");
    if (smp->num_results < 1) return fail(-12, "need num_results >= 1");
    std::lock_guard<std::mutex> lock(g_ws_mutex);
    CUDA_TRY(cudaSetDevice(device));
    const size_t ts = ph->dtype == SCCODA_F64 ? 8 : 4;
    const size_t B = ph->B, C = ph->C, N = ph->N, K = ph->K, D = ph->D, U = ph->Umax, S = smp->num_results;
    const size_t P = D + 2 * D * (K - 1) + K;
    const size_t nsum = SCCODA_NSUM_PARAM * P + SCCODA_NSUM_BETA * D * K + SCCODA_NSUM_SCALAR;
    // one workspace, carved into 256-byte aligned pieces
    struct Piece { const void* host; size_t bytes; size_t off; };
    Piece in[] = {{ph->ey, B * N * K * ts, 0}, {ph->eycst, B * N * K * ts, 0}, {ph->eidx, B * N * K * 4, 0},
                  {ph->colperm, B * K * 4, 0}, {ph->nbig, B * 4, 0}, {ph->ntot, B * N * ts, 0}, {ph->xu, B * U * D * ts, 0},
                  {ph->grp, B * N * 4, 0}, {ph->rstart, B * (U + 1) * 4, 0}, {ph->U, B * 4, 0}, {ph->ref, B * 4, 0},
                  {ph->lconst, B * 8, 0}, {init, B * C * P * ts, 0}};
    const size_t NP = U * (K + 1), NIm = ph->grouped ? (size_t)ph->NImax : 0, NOm = ph->grouped ? (size_t)ph->NOmax : 0,
                 NGm = ph->grouped ? (size_t)ph->NGmax : 0;
    Piece gin[] = {{ph->NI, B * 4, 0}, {ph->iy, B * SCCODA_ITEM_R * NIm * ts, 0}, {ph->ipair, B * NIm * 2, 0},
                   {ph->ipos, B * NIm * 2, 0}, {ph->inreal, B * NIm, 0}, {ph->pcoef, B * NP * 6 * ts, 0}, {ph->NG, B * 4, 0},
                   {ph->oy, B * NOm * ts, 0}, {ph->opair, B * NGm * 2, 0}, {ph->opos, B * NGm * 2, 0},
                   {ph->odesc, B * NGm * 4, 0}};
    size_t off = 0;
    for (auto& p : in) {
        if (!p.host) return fail(-7, "problem has a NULL array");
        p.off = off;
        off += align256(p.bytes);
    }
    if (ph->grouped)
        for (auto& p : gin) {
            if (!p.host) return fail(-7, "grouped problem has a NULL pair/item array");
            p.off = off;
            off += align256(p.bytes);
        }
    const size_t draws_bytes = B * C * S * P * ts, stats_bytes = B * C * S * SCCODA_NSTAT * ts, sum_bytes = B * C * nsum * 8;
    const size_t o_draws = off; off += align256(draws_bytes);
    const size_t o_stats = off; off += align256(stats_bytes);
    const size_t o_sum = off; off += align256(sum_bytes);
    void* base = nullptr;
    int rc = workspace_get(device, off, &base);
    if (rc) return rc;
    char* w = static_cast<char*>(base);
    cudaStream_t st;
    CUDA_TRY(cudaStreamCreate(&st));
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    do {
        bool ok = true;
        for (auto& p : in)
            ok = ok && cudaMemcpyAsync(w + p.off, p.host, p.bytes, cudaMemcpyHostToDevice, st) == cudaSuccess;
        if (ph->grouped)
            for (auto& p : gin)
                ok = ok && cudaMemcpyAsync(w + p.off, p.host, p.bytes, cudaMemcpyHostToDevice, st) == cudaSuccess;
        if (!ok) { rc = fail(-43, "H2D copy failed: %s", cudaGetErrorString(cudaGetLastError())); break; }
        sccoda_problem pd = *ph;
        if (ph->grouped) {
            pd.NI = (const int32_t*)(w + gin[0].off); pd.iy = w + gin[1].off; pd.ipair = (const uint16_t*)(w + gin[2].off);
            pd.ipos = (const uint16_t*)(w + gin[3].off); pd.inreal = (const uint8_t*)(w + gin[4].off);
            pd.pcoef = w + gin[5].off; pd.NG = (const int32_t*)(w + gin[6].off);
            pd.oy = w + gin[7].off; pd.opair = (const uint16_t*)(w + gin[8].off); pd.opos = (const uint16_t*)(w + gin[9].off);
            pd.odesc = (const uint32_t*)(w + gin[10].off);
        }
        pd.ey = w + in[0].off; pd.eycst = w + in[1].off; pd.eidx = (const uint32_t*)(w + in[2].off);
        pd.colperm = (const int32_t*)(w + in[3].off); pd.nbig = (const int32_t*)(w + in[4].off);
        pd.ntot = w + in[5].off; pd.xu = w + in[6].off; pd.grp = (const int32_t*)(w + in[7].off);
        pd.rstart = (const int32_t*)(w + in[8].off); pd.U = (const int32_t*)(w + in[9].off);
        pd.ref = (const int32_t*)(w + in[10].off); pd.lconst = (const double*)(w + in[11].off);
        sccoda_sampler s2 = *smp;
        s2.step_begin = 0; s2.step_end = 0; s2.store_draws = 1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0, st);
        rc = sccoda_sample(&pd, &s2, w + in[12].off, w + o_draws, w + o_stats, nullptr, st);
        if (rc) break;
        cudaEventRecord(e1, st);
        if (summary) {
            rc = sccoda_summarize(&pd, (int32_t)S, summary_skip, w + o_draws, w + o_stats, (double*)(w + o_sum), st);
            if (rc) break;
            cudaMemcpyAsync(summary, w + o_sum, sum_bytes, cudaMemcpyDeviceToHost, st);
        }
        if (draws) cudaMemcpyAsync(draws, w + o_draws, draws_bytes, cudaMemcpyDeviceToHost, st);
        if (stats) cudaMemcpyAsync(stats, w + o_stats, stats_bytes, cudaMemcpyDeviceToHost, st);
        cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { rc = fail(-44, "kernel/copy failed: %s", cudaGetErrorString(e)); break; }
        if (kernel_ms) cudaEventElapsedTime(kernel_ms, e0, e1);
    } while (0);
    if (rc) cudaStreamSynchronize(st);   // error path: nothing may still be copying into the cached workspace
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    cudaStreamDestroy(st);
    return rc;
}

int sccoda_sample_raw_host(int32_t B, int32_t N, int32_t K, int32_t D, int32_t C, const double* x, const double* y,
                           const int32_t* ref, int32_t dtype, const sccoda_sampler* smp, const void* init, void* draws,
                           void* stats, double* summary, int32_t summary_skip, int32_t device, float* timings_ms) {
    using clk = std::chrono::steady_clock;
    const auto t0 = clk::now();
    sccoda_packed* pk = nullptr;
    int rc = sccoda_pack(B, N, K, D, x, y, ref, dtype, 1, -1, 0, &pk);
    if (rc) return rc;
    const auto t1 = clk::now();
    sccoda_problem prob;
    sccoda_packed_problem(pk, C, &prob);
    float kernel_ms = 0.0f;
    rc = sccoda_sample_host(&prob, smp, init, draws, stats, summary, summary_skip, device, &kernel_ms);
    sccoda_packed_free(pk);
    const auto t2 = clk::now();
    if (timings_ms) {
        timings_ms[0] = std::chrono::duration<float, std::milli>(t1 - t0).count();
        timings_ms[1] = kernel_ms;
        timings_ms[2] = std::chrono::duration<float, std::milli>(t2 - t0).count();
    }
    return rc;
}

}  // extern "C"
