/*
 * sccoda_b200 — C ABI of the B200-native MCMC engine for scCODA's Dirichlet-multinomial model.
 *
 * This is the drop-in boundary for the reference's sampler path.  Each entry point names the reference
 * interface it replaces (paths relative to the upstream repository theislab/scCODA):
 *
 *   sccoda_logp_grad      scCODAModel.target_log_prob_fn + TF autodiff      sccoda/model/scCODA_model.py:756-760
 *   sccoda_sample         CompositionalModel.sampling -> tfp.mcmc.sample_chain
 *                         with the kernel stacks of sample_hmc / sample_hmc_da / sample_nuts
 *                                                                            sccoda/model/scCODA_model.py:152-161,
 *                                                                            :276-315, :382-424, :495-553
 *   sccoda_summarize      the per-draw post-processing of get_y_hat and CAResult.complete_beta_df
 *                                                                            sccoda/model/scCODA_model.py:800-834,
 *                                                                            sccoda/util/result_classes.py:240-258
 *   sccoda_predict        the per-draw concentration / prediction arrays of get_y_hat, on demand
 *                                                                            sccoda/model/scCODA_model.py:821-827
 *   sccoda_sample_host    the same sampling call with HOST buffers (what a Python/ctypes caller that owns
 *                         numpy arrays binds; copies host<->device inside)
 *   sccoda_pack           the input handling of CompositionalModel.__init__ (float64 cast, pseudocount, row sums)
 *                                                                            sccoda/model/scCODA_model.py:90-100
 *   sccoda_sample_raw_host  pack + sample from raw x / y host arrays (the end-to-end call the bench times)
 *
 * Conventions
 *   - plain pointers and sizes only; no framework types.  `stream` is a cudaStream_t passed as void*
 *     (NULL = default stream).  All *device* entry points are asynchronous on `stream` and never
 *     synchronise the host.
 *   - every real-valued array has element type float (dtype == SCCODA_F32) or double (SCCODA_F64).
 *   - return value: 0 on success, negative error code otherwise; sccoda_last_error() gives the text.
 *   - thread-compatible, not thread-safe: one host thread per device context.
 *
 * Packed dataset layout (built by sccoda_b200/_pack.py or by the caller; B datasets of equal shape).
 * Rows are sorted by covariate group; the N*K counts form an "element list" stored column-major with the
 * columns (cell types) sorted by decreasing minimum count: element e = kk*N + n.
 *   ey      [B,N*K]     counts after the 0.5 pseudocount (scCODA_model.py:94-96), element-list order
 *   eycst   [B,N*K]     per-count constant: fp32 -> r(y) if y>=6 else lgamma(y); fp64 -> 0   (see csrc/special.cuh)
 *   eidx    [B,N*K] u32 packed index (u*K + k) | (n << 16): covariate group u, ORIGINAL cell type k, row n
 *   colperm [B,K] i32   sorted column position kk -> original cell type k
 *   nbig    [B] i32     the first nbig elements have counts >= 6 (N * number of columns whose minimum is >= 6).
 *                       With grouped != 0 only the value pass walks the list and the order is free: the packer then
 *                       lists ALL counts >= 6 first (stable) and nbig counts them.
 *   ntot    [B,N]       row sums of y (scCODA_model.py:99-100)
 *   xu      [B,Umax,D]  unique covariate rows (design matrix without intercept, comp_ana.py:73-75)
 *   grp     [B,N] i32   covariate group of each row, non-decreasing
 *   rstart  [B,Umax+1] i32  first row of each group (rstart[U..] = N)
 *   U       [B] i32     number of groups of dataset b (<= Umax)
 *   ref     [B] i32     reference cell type (scCODA_model.py:688)
 *   lconst  [B] f64     all parameter-independent terms of the log-density
 *
 * Pair / item layout of the gradient pass (used when `grouped` != 0; built by sccoda_b200/_pack.py: build_items,
 * consumed by csrc/grouped.cuh).  A pair is (covariate group u, cell type k), index u*K + k, or the row totals of a
 * group ("total pair", index Umax*K + u); NP = Umax*(K+1).  Counts >= 6 of a pair are stored SCCODA_ITEM_R at a
 * time in items (padded with 6); counts in {1..5} are folded into the pair's rational; the remaining counts below 6
 * (the 0.5 pseudocount, non-integer data) are "odd" counts, grouped per pair.  Every item and every odd group owns
 * one of the NF result slots of its pair (slot pair*NF + t; padding entries point at the spare slot NP*NF).
 *   NI      [B] i32     number of items of dataset b (<= NImax)
 *   iy      [B,SCCODA_ITEM_R,NImax]  counts of the items (slot-major, so that consecutive items are contiguous)
 *   ipair   [B,NImax] u16    pair of each item
 *   ipos    [B,NImax] u16    result slot of each item
 *   inreal  [B,NImax] u8     how many of the SCCODA_ITEM_R counts of the item are data (the value pass walks the items
 *                            for the counts >= 6 and the tail [nbig, N*K) of the element list for the rest)
 *   pcoef   [B,NP,6]    numerators (a0,b0,a1,b1,a2,b2) of  sum_i W_i/(c+i) = (a0 c+b0)/(c(c+5)) + (a1 c+b1)/((c+1)(c+4))
 *                       + (a2 c+b2)/((c+2)(c+3)),  W_i = #counts>=6 + #odd + #{counts in {1..5} that exceed i}
 *   NG      [B] i32     number of odd groups of dataset b (<= NGmax)
 *   oy      [B,NOmax]   odd counts, the counts of one group consecutive
 *   opair, opos [B,NGmax] u16   pair and result slot of each odd group
 *   odesc   [B,NGmax] u32       first odd count | n_odd << 16
 *
 * State vector (P = D + 2 D (K-1) + K), order of scCODA_model.py:692-693:
 *   [ sigma_d (D) | b_offset (D,K-1) | ind_raw (D,K-1) | alpha (K) ]
 */
#ifndef SCCODA_B200_H
#define SCCODA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCCODA_F32 0
#define SCCODA_F64 1

#define SCCODA_ITEM_R 5 /* counts per item of the pair / item layout */

#define SCCODA_METHOD_HMC 0    /* HMC + SimpleStepSizeAdaptation          (sample_hmc,    :276-289) */
#define SCCODA_METHOD_HMC_DA 1 /* HMC + DualAveragingStepSizeAdaptation   (sample_hmc_da, :382-395) */
#define SCCODA_METHOD_NUTS 2   /* NUTS + DualAveragingStepSizeAdaptation  (sample_nuts,   :495-519) */

/* per-draw statistics, stats[B,C,S,SCCODA_NSTAT] (trace_fn of :296-311, :402-420, :526-550) */
#define SCCODA_NSTAT 8
#define SCCODA_ST_TARGET_LOG_PROB 0
#define SCCODA_ST_LOG_ACCEPT_RATIO 1
#define SCCODA_ST_IS_ACCEPTED 2
#define SCCODA_ST_STEP_SIZE 3
#define SCCODA_ST_DIVERGING 4
#define SCCODA_ST_LEAPFROGS_TAKEN 5
#define SCCODA_ST_ENERGY 6
#define SCCODA_ST_REACH_MAX_DEPTH 7

/* carry[B,C,P+SCCODA_NCARRY] f64: chain state for chunked launches (resume points): the P parameters, then
 * [P+0] step size, [P+1] dual-averaging error sum, [P+2] log averaged step, [P+3] adaptation step count,
 * [P+4] log-density of the state; [P+5], [P+6] are diagnostics the case/control HMC kernel fills (run time of the
 * chain inside the launch in ns, SM it ran on; not read on resume), [P+7] spare */
#define SCCODA_NCARRY 8

/* summary[B,C,SCCODA_NSUM_PARAM*P + SCCODA_NSUM_BETA*D*K + SCCODA_NSUM_SCALAR] f64 — see sccoda_summarize */
#define SCCODA_NSUM_PARAM 2  /* mean, M2 (sum of squared deviations) of every state coordinate */
#define SCCODA_NSUM_BETA 4   /* per (d,k): inclusion count (|beta|>1e-3), sum of included beta, mean, M2 of beta */
#define SCCODA_NSUM_SCALAR 6 /* over ALL traced steps (scCODA_model.py:218-224): draws used after skip, accepted
                                steps, sum exp(min(log_accept_ratio,0)), leapfrogs, diverging steps, last step size */

typedef struct sccoda_problem {
    int32_t B, C;          /* datasets, chains per dataset */
    int32_t N, K, D, Umax; /* samples, cell types, covariates, max covariate groups */
    int32_t dtype;         /* SCCODA_F32 / SCCODA_F64 */
    int32_t nbig_min;      /* min over datasets of nbig[b] (host copy; sizes the shared-memory fallback loop) */
    int32_t grouped;       /* 1: gradient by pairs / items (few covariate groups); 0: by single counts */
    int32_t NImax, NOmax;  /* item / odd-count slots per dataset (grouped layout; 0 when unused) */
    int32_t NGmax, NF;     /* odd-group slots per dataset; result slots per pair (power of two >= 2) */
    int32_t NT, NOT;       /* largest number of items / of odd counts of one pair (over all datasets) */
    int32_t reserved0;
    const void* ey;
    const void* eycst;
    const uint32_t* eidx;
    const int32_t* colperm;
    const int32_t* nbig;
    const void* ntot;
    const void* xu;
    const int32_t* grp;
    const int32_t* rstart;
    const int32_t* U;
    const int32_t* ref;
    const double* lconst;
    /* pair / item layout; may all be NULL when grouped == 0 */
    const int32_t* NI;
    const void* iy;
    const uint16_t* ipair;
    const uint16_t* ipos;
    const uint8_t* inreal;
    const void* pcoef;
    const int32_t* NG;
    const void* oy;
    const uint16_t* opair;
    const uint16_t* opos;
    const uint32_t* odesc;
} sccoda_problem;

typedef struct sccoda_sampler {
    int32_t method;         /* SCCODA_METHOD_* */
    int32_t num_results;    /* traced draws S  (reference default 20000 / 10000) */
    int32_t num_burnin;     /* untraced leading steps (reference default 5000)   */
    int32_t num_adapt;      /* adaptation steps (reference: int(0.8*num_burnin)) */
    int32_t num_leapfrog;   /* HMC leapfrog steps (reference default 10)         */
    int32_t max_tree_depth; /* NUTS (reference default 10)                       */
    int32_t step_begin;     /* chunking: run steps [step_begin, step_end) of num_burnin+num_results */
    int32_t step_end;       /* 0 => run to the end */
    double step_size;       /* initial step size (reference default 0.01) */
    double target_accept;   /* reference: 0.75 */
    uint64_t seed;
    uint32_t chain_id0;     /* global id of chain (b=0,c=0): lets shards on different GPUs keep disjoint streams */
    int32_t warps_per_chain; /* 0 = choose automatically; 1,2,4,8,16 */
    int32_t store_draws;    /* 1 = write draws/stats (default); 0 = only carry (e.g. warm-up chunks) */
    int32_t chains_per_cta; /* 0 = choose automatically; otherwise chain groups per thread block (tuning knob) */
    int32_t generic_only;   /* 1 = never dispatch to a shape-specialised kernel (the case/control kernels); 0 = default */
    int32_t freeze_spike_slab; /* 1 = sigma_d and ind_raw keep their initial values (zero momentum, zero gradient): with
                                  sigma_d = 1 and ind_raw = 1 the density is SimpleModel's (sccoda/model/other_models.py:31-109,
                                  beta = b_offset) up to a constant */
    double reject_conc_above;  /* Concentration guard.  A state with some c_uk = exp(alpha_k + x_u . beta_k) above this limit
                                  gets a NaN log-density and is therefore rejected.  0 = default of the arithmetic type
                                  (1e6 in fp32: the validated range of the fp32 lgamma forms; 1e15 in fp64: where the
                                  reference's own fp64 differences lgamma(c+y) - lgamma(c), scCODA_model.py:746-751, lose
                                  every digit and evaluate to the bare multinomial constant — a spurious mode);
                                  negative or +inf = no guard (with SCCODA_F64 this is the reference's arithmetic, trap
                                  included); any other positive value = that limit.  DESIGN.md §Numerics. */
} sccoda_sampler;

/* Library / device info. */
int sccoda_version(void);
const char* sccoda_last_error(void);

/* Value and gradient of the log-density at theta[B,C,P] -> logp[B,C] (f64), grad[B,C,P] (dtype). */
int sccoda_logp_grad(const sccoda_problem* prob, const void* theta, double* logp, void* grad, void* stream);

/* The same evaluation through the device functions of a given sampler kernel family: kernel_kind as reported by
 * sccoda_kernel_kind (0, 1, 2 = the generic kernels; 3, 4 = the case/control kernels' cc_grad / cc_value; 5, 6 = the
 * multi-group kernels' mg_grad / mg_value) — so that value and gradient of the kernels a sweep actually runs can be
 * compared with the reference density at identical states.  Fails (-25) when the problem does not have the shape that
 * family takes.  reject_conc_above: as in sccoda_sampler. */
int sccoda_logp_grad_as(const sccoda_problem* prob, int32_t kernel_kind, double reject_conc_above, const void* theta,
                        double* logp, void* grad, void* stream);

/* Run the sampler.  init[B,C,P]; draws[B,C,S,P]; stats[B,C,S,SCCODA_NSTAT]; carry[B,C,P+SCCODA_NCARRY] f64
 * (may be NULL when step_begin==0 and the run is not chunked). All device pointers. */
int sccoda_sample(const sccoda_problem* prob, const sccoda_sampler* smp, const void* init, void* draws, void* stats,
                  double* carry, void* stream);

/* Per-chain posterior summaries straight from device-resident draws (skip = python-side second burn-in,
 * scCODA_model.py:205-227).  summary[B,C,*] f64, layout above. */
int sccoda_summarize(const sccoda_problem* prob, int32_t num_results, int32_t skip, const void* draws,
                     const void* stats, double* summary, void* stream);

/* Posterior-predictive quantities per stored draw, straight from device-resident draws — replaces the per-draw
 * python loops of get_y_hat (scCODA_model.py:806-827).  For the draws s in [draw_begin, draw_end) of every chain:
 *   concentration[b,c,s-draw_begin,n,k] = exp(alpha_k + x_n . beta_k),  beta = sigmoid(50 ind_raw) sigma_d b_offset
 *   prediction   [b,c,s-draw_begin,n,k] = n_total_n * concentration / sum_k concentration   (DirichletMultinomial mean)
 * Both outputs have the problem's dtype and shape [B,C,draw_end-draw_begin,N,K]; either may be NULL.
 * row_order [B,N] (device, may be NULL): output row of each packed row — the packed rows are sorted by covariate
 * group, row_order restores the caller's sample order (NULL: rows stay in packed order). */
int sccoda_predict(const sccoda_problem* prob, const int32_t* row_order, int32_t num_results, int32_t draw_begin,
                   int32_t draw_end, const void* draws, void* concentration, void* prediction, void* stream);

/* Launch geometry the sampler would use (for planning / tests): warps per chain, chains per CTA, CTAs, bytes of dynamic
 * shared memory per CTA.  One warp per chain on the lane-per-cell-type kernels; HMC launches with more than ~8 chains
 * per SM that fill whole waves take the wide geometry — one CTA per SM holding a balanced share of ALL chains (up to 28
 * for the case/control kernel) and every dataset they belong to, its warps meeting at a barrier every 16 transitions;
 * an explicit sccoda_sampler::chains_per_cta keeps one dataset per CTA.  Chains do not depend on the geometry. */
int sccoda_launch_plan(const sccoda_problem* prob, const sccoda_sampler* smp, int32_t* warps_per_chain,
                       int32_t* chains_per_cta, int32_t* grid, int32_t* smem_bytes);

/* Which kernel sccoda_sample would launch: 1 = generic HMC, 2 = generic NUTS, 3 / 4 = the case/control HMC / NUTS
 * specialisations (D == 1, Umax == 2 covariate groups, K <= 31, fp32, one warp per chain, grouped layout), 5 = the
 * multi-group HMC kernel (same layout, up to 32 covariate groups; D <= 4 with K <= 31, D <= 3 with two cell types
 * per lane for K <= 62). */
int sccoda_kernel_kind(const sccoda_problem* prob, const sccoda_sampler* smp, int32_t* kind);

/* Host-buffer convenience (the end-to-end path): all pointers are HOST memory; copies in, runs on `device`,
 * copies draws/stats (either may be NULL) and, if summary != NULL, the summaries back.  Synchronous.
 * The device workspace (inputs + draw buffer) is cached across calls; sccoda_release_workspace() frees it. */
int sccoda_sample_host(const sccoda_problem* prob_host, const sccoda_sampler* smp, const void* init, void* draws,
                       void* stats, double* summary, int32_t summary_skip, int32_t device, float* kernel_ms);

int sccoda_release_workspace(void);

/* Host-side packer: raw inputs -> the packed layout above.  Replaces the input handling of
 * CompositionalModel.__init__ (sccoda/model/scCODA_model.py:90-100: float64 cast, zeros -> 0.5 pseudocount BEFORE the
 * row sums) for B datasets at once, threaded over datasets.
 *   x [B,N,D], y [B,N,K] float64 (host), ref [B]; pseudocount != 0 applies the 0 -> 0.5 rule;
 *   grouped: -1 = choose the gradient layout (pair/item for the designs the lane-per-cell-type kernels take and whenever
 *   the item padding stays below 60 % of the counts, as long as the library's planner says it fits), 0 / 1 = force;
 *   threads: 0 = all host cores.
 * The handle owns every array; sccoda_packed_problem fills a sccoda_problem with HOST pointers into it (valid until
 * sccoda_packed_free) — pass it to sccoda_sample_host, or copy the arrays to the device for sccoda_sample.
 * Errors: -52 invalid reference index (the reference raises NameError, comp_ana.py:129-130), -55 the forced pair/item
 * layout overflows its 16-bit indices. */
typedef struct sccoda_packed sccoda_packed;
int sccoda_pack(int32_t B, int32_t N, int32_t K, int32_t D, const double* x, const double* y, const int32_t* ref,
                int32_t dtype, int32_t pseudocount, int32_t grouped, int32_t threads, sccoda_packed** out);
int sccoda_packed_problem(const sccoda_packed* pk, int32_t C, sccoda_problem* prob_host);
const int32_t* sccoda_packed_row_order(const sccoda_packed* pk);  /* [B,N] original row of each packed row */
const double* sccoda_packed_y(const sccoda_packed* pk);           /* [B,N,K] counts after the pseudocount, packed row order */
const double* sccoda_packed_logp_const(const sccoda_packed* pk);  /* [B] multinomial-coefficient constant */
void sccoda_packed_free(sccoda_packed* pk);

/* The whole path from RAW host arrays: pack (above) + sccoda_sample_host.  x [B,N,D], y [B,N,K] float64, ref [B],
 * init [B,C,P] (dtype).  timings_ms (may be NULL) receives [0] host time of the packer, [1] device time of the sampler
 * launch, [2] wall time of the whole call. */
int sccoda_sample_raw_host(int32_t B, int32_t N, int32_t K, int32_t D, int32_t C, const double* x, const double* y,
                           const int32_t* ref, int32_t dtype, const sccoda_sampler* smp, const void* init, void* draws,
                           void* stats, double* summary, int32_t summary_skip, int32_t device, float* timings_ms);

/* Measurement utility for the bench: FP32 FFMA peak (TFLOP/s, 2 flop per FMA) and MUFU peak (T op/s) of the
 * current device, from two register-resident micro-kernels (best of 5). Synchronises `stream`. */
int sccoda_measure_peaks(float* fp32_tflops, float* mufu_tops, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SCCODA_B200_H */
