/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Not product code.
 *
 * Plain-C float64 restatement of the hot path of theislab/scCODA (the same algorithm as
 * oracle/np_oracle.py, which is the line-by-line cited restatement; this file exists so that the
 * statistical parity tests and the CPU baseline run in seconds instead of minutes).
 *
 *   log-density      sccoda/model/scCODA_model.py:702-760   (JointDistributionCoroutine.log_prob)
 *   inputs           sccoda/model/scCODA_model.py:90-100    (pseudocount, n_total) — done by the caller
 *   sample_hmc       sccoda/model/scCODA_model.py:276-315   HMC + SimpleStepSizeAdaptation      [TFP]
 *   sample_hmc_da    sccoda/model/scCODA_model.py:382-424   HMC + DualAveragingStepSizeAdaptation [TFP]
 *   sample_nuts      sccoda/model/scCODA_model.py:495-553   NUTS + DualAveragingStepSizeAdaptation [TFP]
 *   sample_chain     sccoda/model/scCODA_model.py:152-161   num_burnin+num_results steps, last num_results traced
 *
 * [TFP] = tensorflow-probability (>=0.16, requirements.txt:5-6), an un-vendored dependency that cannot
 * be installed here; its published algorithms are restated as in SURVEY.md §8(c).
 * Parity status: log-density/gradient "parity unpinned" by the reference's own tests (it has none for
 * them and TF cannot run here); pinned against scipy/torch-autograd and np_oracle.py in tests/test_oracle.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may load this.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define LOG_2PI 1.8378770664093454835606594728112
#define LOG_2_OVER_PI (-0.45158270528945486472619522989488)
#define NSTAT 8
enum { ST_LP = 0, ST_LAR, ST_ACC, ST_EPS, ST_DIV, ST_LEAP, ST_ENERGY, ST_MAXDEPTH };
enum { METHOD_HMC = 0, METHOD_HMC_DA = 1, METHOD_NUTS = 2 };
enum { STREAM_MOMENTUM = 0, STREAM_ACCEPT = 1, STREAM_NUTS_DIR = 2, STREAM_NUTS_LEAF = 3, STREAM_NUTS_TOP = 4 };

/* ------------------------------------------------------------------ RNG contract (Philox4x32-10) */
static void philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                       uint32_t out[4]) {
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
static inline double u01(uint32_t w) { return ((double)(w >> 8) + 0.5) * (1.0 / 16777216.0); }

typedef struct { uint32_t k0, k1, chain; } rng_t;
static double rng_normal(const rng_t *r, uint32_t index, uint32_t step) {
    uint32_t w[4];
    philox4x32(index, step, r->chain, STREAM_MOMENTUM, r->k0, r->k1, w);
    return sqrt(-2.0 * log(u01(w[0]))) * cos(2.0 * M_PI * u01(w[1]));
}
static double rng_uniform(const rng_t *r, uint32_t index, uint32_t step, uint32_t stream) {
    uint32_t w[4];
    philox4x32(index, step, r->chain, stream, r->k0, r->k1, w);
    return u01(w[0]);
}
static int rng_bit(const rng_t *r, uint32_t index, uint32_t step, uint32_t stream) {
    uint32_t w[4];
    philox4x32(index, step, r->chain, stream, r->k0, r->k1, w);
    return (int)(w[0] & 1u);
}

/* ------------------------------------------------------------------ special functions (fp64) */
static double digamma_pos(double x) {
    /* psi(x), x > 0: recurrence up to x >= 10, then the asymptotic (Bernoulli) series. */
    if (!(x > 0.0)) return (x == 0.0) ? -INFINITY : NAN;
    if (isinf(x)) return x;
    double acc = 0.0;
    while (x < 10.0) { acc -= 1.0 / x; x += 1.0; }
    double w = 1.0 / x, w2 = w * w;
    double ser = w2 * (1.0 / 12.0 - w2 * (1.0 / 120.0 - w2 * (1.0 / 252.0 - w2 * (1.0 / 240.0 - w2 * (1.0 / 132.0
                 - w2 * (691.0 / 32760.0 - w2 * (1.0 / 12.0)))))));
    return acc + log(x) - 0.5 * w - ser;
}
static double lgamma_pos(double x) {
    int sg;
    return lgamma_r(x, &sg);
}

/* ------------------------------------------------------------------ model */
typedef struct {
    int N, K, D, ref, P;
    const double *x, *y, *n_total; /* [N,D], [N,K] (after pseudocount), [N] */
    double logp_const;
} model_t;

/* Optional concentration guard (NOT part of the reference: its density is evaluated for any concentration).  When set,
 * a state with some concentration above the limit gets a NaN value => rejected by the samplers, which is what the
 * device kernels do by default (sccoda_sampler::reject_conc_above).  Default: off = the reference's arithmetic. */
static double g_conc_max = INFINITY;
void oracle_set_reject_conc_above(double limit) { g_conc_max = (limit > 0.0) ? limit : INFINITY; }

typedef struct { double *beta, *ind, *conc, *G, *g; } work_t; /* [D*K], [D*(K-1)], [N*K], [N*K], [D*K] */

static void work_alloc(work_t *w, const model_t *m) {
    w->beta = (double *)calloc((size_t)m->D * m->K, sizeof(double));
    w->ind = (double *)calloc((size_t)m->D * (m->K - 1) + 1, sizeof(double));
    w->conc = (double *)calloc((size_t)m->N * m->K, sizeof(double));
    w->G = (double *)calloc((size_t)m->N * m->K, sizeof(double));
    w->g = (double *)calloc((size_t)m->D * m->K, sizeof(double));
}
static void work_free(work_t *w) { free(w->beta); free(w->ind); free(w->conc); free(w->G); free(w->g); }

static inline double sigmoid_stable(double s) {
    /* equals exp(s)/(1+exp(s)) (scCODA_model.py:724-725) wherever that expression is finite */
    double e = exp(-fabs(s));
    return s >= 0 ? 1.0 / (1.0 + e) : e / (1.0 + e);
}

/* value and/or gradient of the log-density at theta = [sigma_d(D), b_offset(D,K-1), ind_raw(D,K-1), alpha(K)] */
static double logp_grad(const model_t *m, work_t *w, const double *theta, double *grad, int want_value) {
    const int N = m->N, K = m->K, D = m->D, K1 = m->K - 1, ref = m->ref;
    const double *sigma = theta, *boff = theta + D, *iraw = theta + D + D * K1, *alpha = theta + D + 2 * D * K1;
    double lp = 0.0;
    for (int d = 0; d < D; ++d) {
        for (int j = 0; j < K1; ++j) {
            int k = j + (j >= ref);
            double ind = sigmoid_stable(50.0 * iraw[d * K1 + j]);              /* :724-725 */
            w->ind[d * K1 + j] = ind;
            w->beta[d * K + k] = ind * sigma[d] * boff[d * K1 + j];            /* :727-729 */
        }
        w->beta[d * K + ref] = 0.0;                                            /* :732-734 */
    }
    if (want_value) {
        for (int d = 0; d < D; ++d) {                                          /* HalfCauchy(0,1) :703-707 [TFP] */
            if (sigma[d] < 0) lp = -INFINITY;
            else lp += LOG_2_OVER_PI - log1p(sigma[d] * sigma[d]);
        }
        for (int i = 0; i < D * K1; ++i) lp += -0.5 * boff[i] * boff[i] - 0.5 * LOG_2PI;   /* :709-714 */
        for (int i = 0; i < D * K1; ++i) lp += -0.5 * iraw[i] * iraw[i] - 0.5 * LOG_2PI;   /* :717-722 */
        for (int k = 0; k < K; ++k) lp += -0.5 * (alpha[k] / 5.0) * (alpha[k] / 5.0) - log(5.0) - 0.5 * LOG_2PI; /* :736-741 */
        lp += m->logp_const;
    }
    if (grad) memset(w->g, 0, sizeof(double) * D * K);
    if (grad) for (int k = 0; k < K; ++k) grad[D + 2 * D * K1 + k] = -alpha[k] / 25.0;
    for (int n = 0; n < N; ++n) {
        double S = 0.0;
        for (int k = 0; k < K; ++k) {
            double eta = alpha[k];
            for (int d = 0; d < D; ++d) eta += m->x[n * D + d] * w->beta[d * K + k];
            double c = exp(eta);                                               /* :743 */
            w->conc[n * K + k] = c;
            S += c;
            if (want_value && c > g_conc_max) lp = NAN;
        }
        if (want_value) {                                                      /* DirichletMultinomial :746-751 [TFP] */
            for (int k = 0; k < K; ++k)
                lp += lgamma_pos(w->conc[n * K + k] + m->y[n * K + k]) - lgamma_pos(w->conc[n * K + k]);
            lp += lgamma_pos(S) - lgamma_pos(S + m->n_total[n]);
        }
        if (grad) {
            double rowterm = digamma_pos(S + m->n_total[n]) - digamma_pos(S);
            for (int k = 0; k < K; ++k) {
                double c = w->conc[n * K + k];
                double Gnk = c * (digamma_pos(c + m->y[n * K + k]) - digamma_pos(c) - rowterm);
                grad[D + 2 * D * K1 + k] += Gnk;
                for (int d = 0; d < D; ++d) w->g[d * K + k] += m->x[n * D + d] * Gnk;
            }
        }
    }
    if (grad) {
        for (int d = 0; d < D; ++d) {
            double gs = 0.0;
            for (int j = 0; j < K1; ++j) {
                int k = j + (j >= ref), i = d * K1 + j;
                double g = w->g[d * K + k], ind = w->ind[i];
                grad[D + i] = g * ind * sigma[d] - boff[i];
                grad[D + D * K1 + i] = g * sigma[d] * boff[i] * 50.0 * ind * (1.0 - ind) - iraw[i];
                gs += g * ind * boff[i];
            }
            /* HalfCauchy: zero prior gradient for sigma_d < 0 [TFP safe-input where] */
            grad[d] = gs - (sigma[d] >= 0 ? 2.0 * sigma[d] / (1.0 + sigma[d] * sigma[d]) : 0.0);
        }
    }
    return lp;
}

static inline double finite_or_neginf(double v) { return isfinite(v) ? v : -INFINITY; } /* [TFP] safe_sum */
static inline double dot(const double *a, const double *b, int n) { double s = 0; for (int i = 0; i < n; ++i) s += a[i] * b[i]; return s; }
static double logaddexp_(double a, double b) {
    if (a == -INFINITY) return b;
    if (b == -INFINITY) return a;
    double mx = a > b ? a : b;
    return mx + log(exp(a - mx) + exp(b - mx));
}

/* ------------------------------------------------------------------ dual averaging [TFP] */
typedef struct { double mu, H, log_avg, eps, target, decay, gamma, t0; int t, num_adapt; } da_t;
static void da_init(da_t *da, double eps0, int num_adapt, double target) {
    da->mu = log(10.0 * eps0); da->H = 0; da->log_avg = 0; da->eps = eps0; da->target = target;
    da->decay = 0.75; da->gamma = 0.05; da->t0 = 10.0; da->t = 0; da->num_adapt = num_adapt;
}
static void da_update(da_t *da, double log_accept_prob) {
    if (da->t >= da->num_adapt) return;
    da->t += 1;
    double a = (log_accept_prob > -INFINITY) ? exp(log_accept_prob) : 0.0;
    da->H += da->target - a;
    double log_eps = da->mu - da->H * sqrt((double)da->t) / (((double)da->t + da->t0) * da->gamma);
    double eta = pow((double)da->t, -da->decay);
    da->log_avg = eta * log_eps + (1.0 - eta) * da->log_avg;
    da->eps = (da->t < da->num_adapt) ? exp(log_eps) : exp(da->log_avg);
}

/* ------------------------------------------------------------------ one chain */
typedef struct {
    double *q, *p, *g;     /* leapfrog work */
    double *p0;
    /* NUTS */
    double *lq, *lp_, *lg, *rq, *rp, *rg, *cq, *cg, *sq, *sg, *rho, *rho_s, *ck_p, *ck_rho, *span;
} chainbuf_t;

static int popcount_(unsigned v) { int c = 0; while (v) { c += v & 1u; v >>= 1; } return c; }
static int trailing_ones_(unsigned v) { int c = 0; while (v & 1u) { ++c; v >>= 1; } return c; }

static void run_chain(const model_t *m, const double *init, int method, int num_results, int num_burnin,
                      int num_adapt, int L, int max_depth, double eps0, double target, uint64_t seed,
                      uint32_t chain_id, double *draws, double *stats) {
    const int P = m->P;
    work_t w; work_alloc(&w, m);
    rng_t rng = { (uint32_t)(seed & 0xFFFFFFFFu), (uint32_t)(seed >> 32), chain_id };
    double *theta = (double *)malloc(sizeof(double) * P), *g = (double *)malloc(sizeof(double) * P);
    double *q = (double *)malloc(sizeof(double) * P), *p = (double *)malloc(sizeof(double) * P);
    double *gq = (double *)malloc(sizeof(double) * P), *p0 = (double *)malloc(sizeof(double) * P);
    /* NUTS buffers */
    const int nb = 12;
    double *nuts = NULL, *ck = NULL;
    if (method == METHOD_NUTS) {
        nuts = (double *)calloc((size_t)nb * P, sizeof(double));
        ck = (double *)calloc((size_t)2 * (max_depth + 1) * P, sizeof(double));
    }
    memcpy(theta, init, sizeof(double) * P);
    double lp = logp_grad(m, &w, theta, g, 1);
    double eps = eps0;
    da_t da; da_init(&da, eps0, num_adapt, target);
    const int total = num_burnin + num_results;
    for (int t = 0; t < total; ++t) {
        const double eps_used = eps;
        double lar, s_acc, s_div, s_leap, s_energy = NAN, s_maxd = 0.0;
        if (method != METHOD_NUTS) {
            /* [TFP] HamiltonianMonteCarlo.one_step */
            for (int i = 0; i < P; ++i) { p0[i] = rng_normal(&rng, (uint32_t)i, (uint32_t)t); q[i] = theta[i]; }
            for (int i = 0; i < P; ++i) p[i] = p0[i] + 0.5 * eps_used * g[i];
            memcpy(gq, g, sizeof(double) * P);
            double lp_new = lp;
            for (int l = 0; l < L; ++l) {
                for (int i = 0; i < P; ++i) q[i] += eps_used * p[i];
                lp_new = logp_grad(m, &w, q, gq, l == L - 1);
                for (int i = 0; i < P; ++i) p[i] += eps_used * gq[i];
            }
            for (int i = 0; i < P; ++i) p[i] -= 0.5 * eps_used * gq[i];
            lar = finite_or_neginf(lp_new - lp + 0.5 * (dot(p0, p0, P) - dot(p, p, P)));
            double u = rng_uniform(&rng, 0, (uint32_t)t, STREAM_ACCEPT);
            int acc = log(u) < lar;
            if (acc) { memcpy(theta, q, sizeof(double) * P); memcpy(g, gq, sizeof(double) * P); lp = lp_new; }
            s_acc = acc; s_div = (lar < -1000.0); s_leap = L;                   /* diverging: scCODA_model.py:299 */
        } else {
            /* [TFP] NoUTurnSampler.one_step — iterative doubling, multinomial sampling, generalized U-turn */
            double *lq_ = nuts + 0 * P, *lpm = nuts + 1 * P, *lg_ = nuts + 2 * P;
            double *rq_ = nuts + 3 * P, *rpm = nuts + 4 * P, *rg_ = nuts + 5 * P;
            double *cq_ = nuts + 6 * P, *cg_ = nuts + 7 * P, *sq_ = nuts + 8 * P, *sg_ = nuts + 9 * P;
            double *rho = nuts + 10 * P, *rho_s = nuts + 11 * P;
            double *ck_p = ck, *ck_rho = ck + (size_t)(max_depth + 1) * P;
            for (int i = 0; i < P; ++i) p0[i] = rng_normal(&rng, (uint32_t)i, (uint32_t)t);
            const double e0 = lp - 0.5 * dot(p0, p0, P);
            memcpy(lq_, theta, sizeof(double) * P); memcpy(lpm, p0, sizeof(double) * P); memcpy(lg_, g, sizeof(double) * P);
            memcpy(rq_, theta, sizeof(double) * P); memcpy(rpm, p0, sizeof(double) * P); memcpy(rg_, g, sizeof(double) * P);
            memcpy(cq_, theta, sizeof(double) * P); memcpy(cg_, g, sizeof(double) * P);
            memcpy(rho, p0, sizeof(double) * P);
            double l_lp = lp, r_lp = lp, c_lp = lp, c_energy = e0, s_lp = lp, s_en = e0;
            double wsum = 0.0, accept_sum = 0.0;
            int leapfrogs = 0, is_acc = 0, cont = 1, not_div = 1, depth = 0;
            uint32_t leaf_serial = 0;
            while (depth < max_depth && cont) {
                const int forward = rng_bit(&rng, (uint32_t)depth, (uint32_t)t, STREAM_NUTS_DIR);
                double *eq = forward ? rq_ : lq_, *ep = forward ? rpm : lpm, *eg = forward ? rg_ : lg_;
                double e_lp = forward ? r_lp : l_lp;
                const double se = forward ? eps_used : -eps_used;
                memcpy(q, eq, sizeof(double) * P); memcpy(p, ep, sizeof(double) * P); memcpy(gq, eg, sizeof(double) * P);
                double lq = e_lp;
                const unsigned n_leaves = 1u << depth;
                double w_s = -INFINITY;
                memset(rho_s, 0, sizeof(double) * P);
                int alive = 1;
                for (unsigned i = 0; i < n_leaves && alive; ++i) {
                    for (int j = 0; j < P; ++j) { p[j] += 0.5 * se * gq[j]; q[j] += se * p[j]; }
                    lq = logp_grad(m, &w, q, gq, 1);
                    for (int j = 0; j < P; ++j) { p[j] += 0.5 * se * gq[j]; rho_s[j] += p[j]; }
                    leapfrogs += 1;
                    const int idx_max = popcount_(i >> 1);
                    int no_uturn = 1;
                    if ((i & 1u) == 0) {
                        memcpy(ck_p + (size_t)idx_max * P, p, sizeof(double) * P);
                        memcpy(ck_rho + (size_t)idx_max * P, rho_s, sizeof(double) * P);
                    } else {
                        const int n_chk = trailing_ones_(i);
                        for (int s = idx_max; s > idx_max - n_chk && no_uturn; --s) {
                            const double *cp = ck_p + (size_t)s * P, *cr = ck_rho + (size_t)s * P;
                            double d1 = 0, d2 = 0;
                            for (int j = 0; j < P; ++j) { double sp = rho_s[j] - cr[j] + cp[j]; d1 += sp * cp[j]; d2 += sp * p[j]; }
                            if (!(d1 >= 0.0 && d2 >= 0.0)) no_uturn = 0;
                        }
                    }
                    double energy = lq - 0.5 * dot(p, p, P);
                    if (isnan(energy)) energy = -INFINITY;
                    const double delta = energy - e0;
                    const int divergent = !(-delta < 1000.0);
                    const double w_new = logaddexp_(w_s, delta);
                    const double u = rng_uniform(&rng, leaf_serial++, (uint32_t)t, STREAM_NUTS_LEAF);
                    const double thresh = (w_new > -INFINITY) ? delta - w_new : NAN;
                    if (log1p(-u) <= thresh) {
                        memcpy(sq_, q, sizeof(double) * P); memcpy(sg_, gq, sizeof(double) * P); s_lp = lq; s_en = energy;
                    }
                    w_s = w_new;
                    accept_sum += exp(delta < 0.0 ? delta : 0.0);
                    if (divergent) not_div = 0;
                    alive = (!divergent) && no_uturn;
                }
                const double tree_w = alive ? w_s : -INFINITY;
                double thresh = tree_w - wsum;
                if (isnan(thresh)) thresh = 0.0;
                const double u = rng_uniform(&rng, (uint32_t)depth, (uint32_t)t, STREAM_NUTS_TOP);
                if ((log1p(-u) <= thresh) && alive) {
                    memcpy(cq_, sq_, sizeof(double) * P); memcpy(cg_, sg_, sizeof(double) * P);
                    c_lp = s_lp; c_energy = s_en; is_acc = 1;
                }
                wsum = logaddexp_(wsum, tree_w);
                for (int j = 0; j < P; ++j) rho[j] += rho_s[j];
                memcpy(eq, q, sizeof(double) * P); memcpy(ep, p, sizeof(double) * P); memcpy(eg, gq, sizeof(double) * P);
                if (forward) r_lp = lq; else l_lp = lq;
                const int no_uturn_traj = (dot(rho, lpm, P) >= 0.0) && (dot(rho, rpm, P) >= 0.0);
                cont = alive && no_uturn_traj;
                depth += 1;
            }
            memcpy(theta, cq_, sizeof(double) * P); memcpy(g, cg_, sizeof(double) * P); lp = c_lp;
            lar = accept_sum > 0 ? log(accept_sum / leapfrogs) : -INFINITY;
            s_acc = is_acc; s_div = !not_div; s_leap = leapfrogs; s_energy = c_energy; s_maxd = cont;
        }
        /* step-size adaptation */
        double traced_eps = eps_used;
        if (method == METHOD_HMC) {                                            /* [TFP] SimpleStepSizeAdaptation, rate 0.01 */
            if (t < num_adapt) eps = ((lar < 0 ? lar : 0.0) > log(target)) ? eps * 1.01 : eps / 1.01;
        } else {
            da_update(&da, lar < 0 ? lar : 0.0);
            eps = da.eps;
            if (method == METHOD_HMC_DA) traced_eps = exp(da.log_avg);         /* scCODA_model.py:408,419 */
        }
        const int r = t - num_burnin;
        if (r >= 0) {
            memcpy(draws + (size_t)r * P, theta, sizeof(double) * P);
            double *st = stats + (size_t)r * NSTAT;
            st[ST_LP] = lp; st[ST_LAR] = lar; st[ST_ACC] = s_acc; st[ST_EPS] = traced_eps; st[ST_DIV] = s_div;
            st[ST_LEAP] = s_leap; st[ST_ENERGY] = s_energy; st[ST_MAXDEPTH] = s_maxd;
        }
    }
    free(theta); free(g); free(q); free(p); free(gq); free(p0); free(nuts); free(ck);
    work_free(&w);
}

/* ------------------------------------------------------------------ exported entry points (ctypes) */
int oracle_logp_grad(int N, int K, int D, int ref, const double *x, const double *y, const double *n_total,
                     double logp_const, const double *theta, double *logp_out, double *grad_out) {
    model_t m = { N, K, D, ref, D + 2 * D * (K - 1) + K, x, y, n_total, logp_const };
    work_t w; work_alloc(&w, &m);
    *logp_out = logp_grad(&m, &w, theta, grad_out, 1);
    work_free(&w);
    return 0;
}

/* B datasets x C chains; draws [B,C,S,P], stats [B,C,S,NSTAT]; chain ids = chain_id0 + b*C + c.
 * Returns the number of gradient evaluations performed (summed over chains) via *grad_evals. */
int oracle_sample(int B, int C, int N, int K, int D, const int *ref, const double *x, const double *y,
                  const double *n_total, const double *logp_const, const double *init, int method,
                  int num_results, int num_burnin, int num_adapt, int L, int max_depth, double eps0,
                  double target, uint64_t seed, uint32_t chain_id0, double *draws, double *stats, int n_threads) {
    const int P = D + 2 * D * (K - 1) + K;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (int bc = 0; bc < B * C; ++bc) {
        const int b = bc / C;
        model_t m = { N, K, D, ref[b], P, x + (size_t)b * N * D, y + (size_t)b * N * K, n_total + (size_t)b * N,
                      logp_const[b] };
        run_chain(&m, init + (size_t)bc * P, method, num_results, num_burnin, num_adapt, L, max_depth, eps0, target,
                  seed, chain_id0 + (uint32_t)bc, draws + (size_t)bc * num_results * P,
                  stats + (size_t)bc * num_results * NSTAT);
    }
    return 0;
}

int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
