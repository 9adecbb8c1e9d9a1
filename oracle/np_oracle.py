"""
ORACLE — TEST INFRASTRUCTURE ONLY.  Not product code.

CPU restatement (numpy / scipy, float64) of the hot path of theislab/scCODA:
the Dirichlet-multinomial spike-and-slab log-density, its gradient and the
TensorFlow-Probability samplers the reference drives (HMC + simple step-size
adaptation, HMC + dual averaging, NUTS + dual averaging).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module.  The product package
``sccoda_b200`` never does.

Parity status
-------------
* TensorFlow / TensorFlow-Probability are NOT installable in this environment
  (no wheels, no network), so the reference graph itself cannot be executed.
  The log-density / gradient are therefore **parity unpinned by the reference's
  own tests** (the reference has no log-prob unit test at all).
  They are pinned instead against
    - an independent torch-autograd float64 evaluation of the same density
      (tests/test_oracle.py), and
    - the reference's own end-to-end assertions (tests/unit_tests.py:164-265:
      expected-sample counts within +-30 cells; Condition2 has no credible
      effect) and the captured tutorial outputs, which the
      sequential samplers below reproduce.
* The data generator restatement IS pinned by the reference's golden vector
  (tests/unit_tests.py:44-76) — see oracle/np_datagen.py.

All ``file:line`` citations are relative to /root/reference/.
[TFP] marks semantics of tensorflow-probability (>=0.16), an un-vendored pip
dependency of the reference (requirements.txt:5-6), restated from its
published algorithm.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
from scipy.special import digamma, gammaln

LOG_2PI = math.log(2.0 * math.pi)
LOG_2_OVER_PI = math.log(2.0 / math.pi)

# --------------------------------------------------------------------------------------
# RNG contract shared with the CUDA kernels (sccoda_b200/csrc/philox.cuh).
# The reference uses TF's stateful RNG, which cannot be reproduced; the streams below are
# OUR contract.  Philox4x32-10 (Salmon et al. 2011), key = 64-bit seed,
# counter = (index, step, chain, stream).
# --------------------------------------------------------------------------------------
STREAM_MOMENTUM = 0   # index = flat parameter index
STREAM_ACCEPT = 1     # index = 0
STREAM_NUTS_DIR = 2   # index = tree depth
STREAM_NUTS_LEAF = 3  # index = leaf serial number within the transition
STREAM_NUTS_TOP = 4   # index = tree depth

_M0, _M1 = 0xD2511F53, 0xCD9E8D57
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = 0xFFFFFFFF


def philox4x32(c0: int, c1: int, c2: int, c3: int, k0: int, k1: int):
    """Philox4x32-10 block function on python ints (bit-exact contract)."""
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = (p0 >> 32) & _MASK, p0 & _MASK
        hi1, lo1 = (p1 >> 32) & _MASK, p1 & _MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & _MASK, lo1, (hi0 ^ c3 ^ k1) & _MASK, lo0
        k0 = (k0 + _W0) & _MASK
        k1 = (k1 + _W1) & _MASK
    return c0, c1, c2, c3


def u01(word: int) -> float:
    """24-bit uniform in (0,1): exactly representable in fp32 and fp64."""
    return ((word >> 8) + 0.5) * (1.0 / 16777216.0)


class FastRNG:
    """numpy Generator with the ChainRNG interface — NOT the device contract; used only when the oracle is
    *timed* as the CPU baseline, so that the pure-python Philox does not inflate the baseline's run time."""

    def __init__(self, seed: int, chain: int):
        self.g = np.random.default_rng([seed, chain])

    def normals(self, n: int, step: int) -> np.ndarray:
        return self.g.standard_normal(n)

    def uniform(self, index: int, step: int, stream: int) -> float:
        return float(self.g.random())

    def bit(self, index: int, step: int, stream: int) -> int:
        return int(self.g.integers(2))


class ChainRNG:
    """Counter-based draws for one chain; mirrors the device functions one to one."""

    def __init__(self, seed: int, chain: int):
        self.k0 = seed & _MASK
        self.k1 = (seed >> 32) & _MASK
        self.chain = chain & _MASK

    def words(self, index: int, step: int, stream: int):
        return philox4x32(index & _MASK, step & _MASK, self.chain, stream, self.k0, self.k1)

    def normal(self, index: int, step: int) -> float:
        w = self.words(index, step, STREAM_MOMENTUM)
        return math.sqrt(-2.0 * math.log(u01(w[0]))) * math.cos(2.0 * math.pi * u01(w[1]))

    def normals(self, n: int, step: int) -> np.ndarray:
        return np.array([self.normal(i, step) for i in range(n)])

    def uniform(self, index: int, step: int, stream: int) -> float:
        return u01(self.words(index, step, stream)[0])

    def bit(self, index: int, step: int, stream: int) -> int:
        return self.words(index, step, stream)[0] & 1


# --------------------------------------------------------------------------------------
# Model data  (sccoda/model/scCODA_model.py:63-113, 667-770)
# --------------------------------------------------------------------------------------
@dataclass
class ModelData:
    x: np.ndarray          # [N, D] covariates
    y: np.ndarray          # [N, K] counts after pseudocount
    n_total: np.ndarray    # [N]
    ref: int               # reference cell type index
    logp_const: float      # sum_n [lgamma(n_n+1) - sum_k lgamma(y_nk+1)]
    freeze: bool = False   # SimpleModel runs: sigma_d and ind_raw keep their values (zero momentum, zero gradient)
    # NOT part of the reference (its density is evaluated for any concentration): when finite, a state with a
    # concentration above the limit has a NaN value => rejected, as the device kernels do by default
    reject_conc_above: float = float("inf")

    def frozen_mask(self) -> np.ndarray:
        """True for the parameters a SimpleModel run holds fixed (sigma_d and ind_raw)."""
        D, DK1 = self.D, self.D * (self.K - 1)
        mask = np.zeros(self.P, dtype=bool)
        if self.freeze:
            mask[:D] = True
            mask[D + DK1:D + 2 * DK1] = True
        return mask

    @property
    def N(self):
        return self.y.shape[0]

    @property
    def K(self):
        return self.y.shape[1]

    @property
    def D(self):
        return self.x.shape[1]

    @property
    def P(self):
        return self.D + 2 * self.D * (self.K - 1) + self.K


def prepare(x, y, ref: int) -> ModelData:
    """scCODA_model.py:90-113: fp64 cast; zeros -> 0.5 pseudocount BEFORE the row sums."""
    x = np.array(x, dtype=np.float64)
    y = np.array(y, dtype=np.float64)
    if x.ndim == 1:
        x = x[:, None]
    if np.count_nonzero(y) != y.size:                        # :94-96
        y[y == 0] = 0.5
    n_total = y.sum(axis=1)                                  # :99-100
    if x.shape[0] != y.shape[0]:                             # :110-111
        raise ValueError("Wrong input dimensions X[{},:] != y[{},:]".format(x.shape[0], y.shape[0]))
    const = float(np.sum(gammaln(n_total + 1.0)) - np.sum(gammaln(y + 1.0)))
    return ModelData(x=x, y=y, n_total=n_total, ref=int(ref), logp_const=const)


def split(theta: np.ndarray, m: ModelData):
    """Flat state -> (sigma_d[D], b_offset[D,K-1], ind_raw[D,K-1], alpha[K]); order of :692-693, :763-768."""
    D, K = m.D, m.K
    o = 0
    sigma = theta[o:o + D]; o += D
    boff = theta[o:o + D * (K - 1)].reshape(D, K - 1); o += D * (K - 1)
    iraw = theta[o:o + D * (K - 1)].reshape(D, K - 1); o += D * (K - 1)
    alpha = theta[o:o + K]
    return sigma, boff, iraw, alpha


def stable_sigmoid(s):
    """Equal to exp(s)/(1+exp(s)) (scCODA_model.py:724-725) wherever that is finite."""
    s = np.asarray(s, dtype=np.float64)
    e = np.exp(-np.abs(s))
    return np.where(s >= 0, 1.0 / (1.0 + e), e / (1.0 + e))


def _beta_full(sigma, boff, iraw, m: ModelData):
    ind = stable_sigmoid(50.0 * iraw)                         # :724-725
    b_raw = sigma[:, None] * boff                             # :727
    beta_nb = ind * b_raw                                     # :729
    beta = np.insert(beta_nb, m.ref, 0.0, axis=1)             # :732-734
    return ind, b_raw, beta


def logp(theta: np.ndarray, m: ModelData) -> float:
    """target_log_prob_fn (scCODA_model.py:702-760), float64."""
    sigma, boff, iraw, alpha = split(theta, m)
    with np.errstate(all="ignore"):
        # HalfCauchy(0,1) on the UNCONSTRAINED sigma_d (:703-707) [TFP]: -inf for sigma_d < 0
        if np.any(sigma < 0):
            lp_sigma = -np.inf
        else:
            lp_sigma = np.sum(LOG_2_OVER_PI - np.log1p(sigma ** 2))
        lp_boff = np.sum(-0.5 * boff ** 2 - 0.5 * LOG_2PI)    # :709-714
        lp_iraw = np.sum(-0.5 * iraw ** 2 - 0.5 * LOG_2PI)    # :717-722
        lp_alpha = np.sum(-0.5 * (alpha / 5.0) ** 2 - math.log(5.0) - 0.5 * LOG_2PI)  # :736-741
        _, _, beta = _beta_full(sigma, boff, iraw, m)
        conc = np.exp(alpha[None, :] + m.x @ beta)            # :743
        if np.any(conc > m.reject_conc_above):
            return float("nan")
        S = conc.sum(axis=1)
        # DirichletMultinomial [TFP]: lbeta(c+y) - lbeta(c) + log_combinations(n, y)   (:746-751)
        lp_dm = (np.sum(gammaln(conc + m.y) - gammaln(conc))
                 - np.sum(gammaln(S + m.n_total)) + np.sum(gammaln(S)) + m.logp_const)
        return float(lp_sigma + lp_boff + lp_iraw + lp_alpha + lp_dm)


def grad(theta: np.ndarray, m: ModelData) -> np.ndarray:
    """Analytic gradient of logp; what TF autodiff yields for the graph of :702-751."""
    sigma, boff, iraw, alpha = split(theta, m)
    with np.errstate(all="ignore"):
        ind, b_raw, beta = _beta_full(sigma, boff, iraw, m)
        conc = np.exp(alpha[None, :] + m.x @ beta)
        S = conc.sum(axis=1)
        rowterm = digamma(S + m.n_total) - digamma(S)
        G = conc * (digamma(conc + m.y) - digamma(conc) - rowterm[:, None])   # d/d log c_nk
        g_alpha = G.sum(axis=0) - alpha / 25.0
        g_full = m.x.T @ G                                                     # [D, K]
        g = np.delete(g_full, m.ref, axis=1)                                   # only k != ref
        g_boff = g * ind * sigma[:, None] - boff
        # HalfCauchy: zero prior gradient for sigma_d < 0 [TFP safe-input where]
        g_sigma = np.sum(g * ind * boff, axis=1) - np.where(sigma >= 0, 2.0 * sigma / (1.0 + sigma ** 2), 0.0)
        g_iraw = g * sigma[:, None] * boff * 50.0 * ind * (1.0 - ind) - iraw
        out = np.concatenate([g_sigma, g_boff.ravel(), g_iraw.ravel(), g_alpha])
        if m.freeze:
            out[m.frozen_mask()] = 0.0
        return out


def logp_and_grad(theta, m):
    return logp(theta, m), grad(theta, m)


# --------------------------------------------------------------------------------------
# Samplers [TFP]  (driven by scCODA_model.py:115-169, 229-567)
# --------------------------------------------------------------------------------------
METHOD_HMC = 0      # sample_hmc     (:276-315)  HMC + SimpleStepSizeAdaptation
METHOD_HMC_DA = 1   # sample_hmc_da  (:382-424)  HMC + DualAveragingStepSizeAdaptation
METHOD_NUTS = 2     # sample_nuts    (:495-553)  NUTS + DualAveragingStepSizeAdaptation


def _finite_or_neginf(v: float) -> float:
    """[TFP] mcmc_util.safe_sum: a non-finite sum becomes -inf."""
    return v if math.isfinite(v) else -math.inf


class DualAveraging:
    """[TFP] DualAveragingStepSizeAdaptation: gamma=0.05, t0=10, kappa=decay_rate, mu=log(10*eps0)."""

    def __init__(self, eps0, num_adapt, target=0.75, decay=0.75, gamma=0.05, t0=10.0):
        self.mu = math.log(10.0 * eps0)
        self.num_adapt, self.target, self.decay, self.gamma, self.t0 = num_adapt, target, decay, gamma, t0
        self.H = 0.0
        self.log_avg = 0.0
        self.t = 0
        self.eps = eps0

    def update(self, log_accept_prob: float):
        if self.t >= self.num_adapt:       # frozen afterwards
            return
        self.t += 1
        a = math.exp(log_accept_prob) if log_accept_prob > -math.inf else 0.0
        self.H += self.target - a
        log_eps = self.mu - self.H * math.sqrt(self.t) / ((self.t + self.t0) * self.gamma)
        eta = self.t ** (-self.decay)
        self.log_avg = eta * log_eps + (1.0 - eta) * self.log_avg
        self.eps = math.exp(log_eps) if self.t < self.num_adapt else math.exp(self.log_avg)


def hmc_transition(theta, lp, g, eps, L, m, rng: ChainRNG, step):
    """[TFP] HamiltonianMonteCarlo.one_step: leapfrog with L gradient evaluations + Metropolis-Hastings."""
    P = theta.shape[0]
    p0 = rng.normals(P, step)
    if m.freeze:
        p0[m.frozen_mask()] = 0.0
    q = theta.copy()
    p = p0 + 0.5 * eps * g
    gq = g
    for _ in range(L):
        q = q + eps * p
        gq = grad(q, m)
        p = p + eps * gq
    p = p - 0.5 * eps * gq
    lp_new = logp(q, m)
    with np.errstate(all="ignore"):
        lar = _finite_or_neginf(lp_new - lp + 0.5 * (float(p0 @ p0) - float(p @ p)))
    u = rng.uniform(0, step, STREAM_ACCEPT)
    accepted = math.log(u) < lar
    if accepted:
        return q, lp_new, gq, lar, True
    return theta, lp, g, lar, False


def _logaddexp(a, b):
    if a == -math.inf:
        return b
    if b == -math.inf:
        return a
    mx = max(a, b)
    return mx + math.log(math.exp(a - mx) + math.exp(b - mx))


def _popcount(v):
    return bin(v).count("1")


def _trailing_ones(v):
    c = 0
    while v & 1:
        c += 1
        v >>= 1
    return c


def nuts_transition(theta, lp, g, eps, max_depth, m, rng: ChainRNG, step, max_energy_diff=1000.0):
    """[TFP] NoUTurnSampler.one_step: iterative tree doubling, multinomial sampling, generalized U-turn.

    energy := logp - 0.5|p|^2.  Restated per SURVEY §8c.
    """
    P = theta.shape[0]
    p_init = rng.normals(P, step)
    if m.freeze:
        p_init[m.frozen_mask()] = 0.0
    e0 = lp - 0.5 * float(p_init @ p_init)
    # trajectory ends: (q, p, lp, g)
    left = (theta, p_init, lp, g)
    right = (theta, p_init, lp, g)
    cand = (theta, lp, g, e0)
    w = 0.0                       # log weight of the trajectory so far
    rho = p_init.copy()
    accept_sum, leapfrogs = 0.0, 0
    is_accepted, cont, not_div = False, True, True
    leaf_serial = 0
    depth = 0
    ck_p = np.zeros((max_depth + 1, P))
    ck_rho = np.zeros((max_depth + 1, P))
    while depth < max_depth and cont:
        forward = rng.bit(depth, step, STREAM_NUTS_DIR) == 1
        q, p, lq, gq = right if forward else left
        se = eps if forward else -eps
        n_leaves = 1 << depth
        # ---- build subtree ----
        w_s = -math.inf
        rho_s = np.zeros(P)
        cand_s = None
        alive = True
        i = 0
        while i < n_leaves and alive:
            # one leapfrog (unrolled_leapfrog_steps=1)
            ph = p + 0.5 * se * gq
            q = q + se * ph
            lq, gq = logp(q, m), grad(q, m)
            p = ph + 0.5 * se * gq
            rho_s = rho_s + p
            leapfrogs += 1
            # iterative U-turn bookkeeping
            idx_max = _popcount(i >> 1)
            no_uturn = True
            if i % 2 == 0:
                ck_p[idx_max] = p
                ck_rho[idx_max] = rho_s
            else:
                n_chk = _trailing_ones(i)
                for s in range(idx_max, idx_max - n_chk, -1):
                    span = rho_s - ck_rho[s] + ck_p[s]
                    if not (float(span @ ck_p[s]) >= 0.0 and float(span @ p) >= 0.0):
                        no_uturn = False
                        break
            with np.errstate(all="ignore"):
                energy = lq - 0.5 * float(p @ p)
            if math.isnan(energy):
                energy = -math.inf
            delta = energy - e0
            divergent = not (-delta < max_energy_diff)
            w_new = _logaddexp(w_s, delta)
            u = rng.uniform(leaf_serial, step, STREAM_NUTS_LEAF)
            leaf_serial += 1
            thresh = delta - w_new if w_new > -math.inf else math.nan
            if math.log1p(-u) <= thresh:
                cand_s = (q, lq, gq, energy)
            w_s = w_new
            accept_sum += math.exp(min(delta, 0.0))
            if divergent:
                not_div = False
            alive = (not divergent) and no_uturn
            i += 1
        # ---- merge subtree ----
        tree_w = w_s if alive else -math.inf
        thresh = tree_w - w
        if math.isnan(thresh):
            thresh = 0.0
        u = rng.uniform(depth, step, STREAM_NUTS_TOP)
        if (math.log1p(-u) <= thresh) and alive:
            cand = cand_s
            is_accepted = True
        w = _logaddexp(w, tree_w)
        rho = rho + rho_s
        if forward:
            right = (q, p, lq, gq)
        else:
            left = (q, p, lq, gq)
        no_uturn_traj = float(rho @ left[1]) >= 0.0 and float(rho @ right[1]) >= 0.0
        cont = alive and no_uturn_traj
        depth += 1
    lar = math.log(accept_sum / leapfrogs) if accept_sum > 0 else -math.inf
    stats = dict(log_accept_ratio=lar, leapfrogs_taken=leapfrogs, is_accepted=is_accepted,
                 reach_max_depth=cont, diverging=not not_div, energy=cand[3], target_log_prob=cand[1])
    return cand[0], cand[1], cand[2], stats


def sample_chain(m: ModelData, init: np.ndarray, method: int, num_results: int, num_burnin: int,
                 num_adapt: int | None = None, num_leapfrog: int = 10, max_tree_depth: int = 10,
                 step_size: float = 0.01, seed: int = 0, chain: int = 0, target_accept: float = 0.75,
                 fast_rng: bool = False):
    """tfp.mcmc.sample_chain as driven by CompositionalModel.sampling (scCODA_model.py:152-161).

    Runs num_burnin + num_results transitions, traces the last num_results (the python-side second
    burn-in slice of :205-227 is applied by the caller).
    """
    if num_adapt is None:
        num_adapt = int(0.8 * num_burnin)                     # :284-285
    rng = FastRNG(seed, chain) if fast_rng else ChainRNG(seed, chain)
    theta = np.array(init, dtype=np.float64)
    lp, g = logp(theta, m), grad(theta, m)
    P = theta.shape[0]
    draws = np.zeros((num_results, P))
    st = {k: np.zeros(num_results) for k in
          ("target_log_prob", "log_accept_ratio", "is_accepted", "step_size", "diverging",
           "leapfrogs_taken", "energy", "reach_max_depth")}
    eps = step_size
    da = DualAveraging(step_size, num_adapt, target=target_accept) if method != METHOD_HMC else None
    total = num_burnin + num_results
    for t in range(total):
        eps_used = eps
        if method == METHOD_NUTS:
            theta, lp, g, s = nuts_transition(theta, lp, g, eps_used, max_tree_depth, m, rng, t)
            lar = s["log_accept_ratio"]
        else:
            theta, lp, g, lar, acc = hmc_transition(theta, lp, g, eps_used, num_leapfrog, m, rng, t)
            s = dict(log_accept_ratio=lar, is_accepted=acc, diverging=lar < -1000.0,     # :299
                     leapfrogs_taken=num_leapfrog, energy=math.nan, reach_max_depth=False, target_log_prob=lp)
        # step-size adaptation
        if method == METHOD_HMC:
            # [TFP] SimpleStepSizeAdaptation, adaptation_rate=0.01, target 0.75 (:288-289)
            if t < num_adapt:
                eps = eps * 1.01 if min(lar, 0.0) > math.log(target_accept) else eps / 1.01
            traced_eps = eps_used
        else:
            da.update(min(lar, 0.0))
            eps = da.eps
            # sample_hmc_da traces exp(log_averaging_step) (:408,419); sample_nuts traces the step used (:533,547)
            traced_eps = math.exp(da.log_avg) if method == METHOD_HMC_DA else eps_used
        r = t - num_burnin
        if r >= 0:
            draws[r] = theta
            st["target_log_prob"][r] = lp
            st["log_accept_ratio"][r] = lar
            st["is_accepted"][r] = float(s["is_accepted"])
            st["step_size"][r] = traced_eps
            st["diverging"][r] = float(s["diverging"])
            st["leapfrogs_taken"][r] = s["leapfrogs_taken"]
            st["energy"][r] = s["energy"]
            st["reach_max_depth"][r] = float(s["reach_max_depth"])
    return draws, st


# --------------------------------------------------------------------------------------
# SimpleModel (sccoda/model/other_models.py:31-109): Dirichlet-multinomial with Normal priors, no spike-and-slab
# --------------------------------------------------------------------------------------
def simple_theta(b: np.ndarray, alpha: np.ndarray) -> np.ndarray:
    """State of the full model that realises SimpleModel: sigma_d = 1, b_offset = b, ind_raw = 1 (indicator = 1 to
    machine precision), alpha."""
    b = np.asarray(b, dtype=np.float64)
    D = b.shape[0]
    return np.concatenate([np.ones(D), b.ravel(), np.ones(b.size), np.asarray(alpha, dtype=np.float64)])


def simple_logp_offset(D: int, K: int) -> float:
    """logp_full(simple_theta(b, alpha)) - logp_SimpleModel(b, alpha): the priors of the frozen parameters,
    HalfCauchy(0,1) at sigma_d = 1 (:703-707 of scCODA_model.py) and Normal(0,1) at ind_raw = 1 (:717-722)."""
    return D * (LOG_2_OVER_PI - math.log(2.0)) + D * (K - 1) * (-0.5 - 0.5 * LOG_2PI)


def simple_logp(b: np.ndarray, alpha: np.ndarray, m: ModelData) -> float:
    """SimpleModel.target_log_prob_fn (other_models.py:70-104): b ~ N(0,1), alpha ~ N(0,5), beta = b with a zero
    reference column, y ~ DirichletMultinomial(exp(alpha + x beta))."""
    return logp(simple_theta(b, alpha), m) - simple_logp_offset(m.D, m.K)


# --------------------------------------------------------------------------------------
# Post-processing restatements
# --------------------------------------------------------------------------------------
def derived_draws(draws: np.ndarray, m: ModelData):
    """get_y_hat (scCODA_model.py:800-840): ind, b_raw, beta, concentration, prediction per draw + y_hat."""
    S = draws.shape[0]
    D, K = m.D, m.K
    o = 0
    sigma = draws[:, o:o + D].reshape(S, D, 1); o += D
    boff = draws[:, o:o + D * (K - 1)].reshape(S, D, K - 1); o += D * (K - 1)
    iraw = draws[:, o:o + D * (K - 1)].reshape(S, D, K - 1); o += D * (K - 1)
    alpha = draws[:, o:o + K]
    ind = stable_sigmoid(50.0 * iraw)                                    # :806-810
    b_raw = boff * sigma                                                 # :812
    beta_nb = ind * b_raw                                                # :814
    beta = np.insert(beta_nb, m.ref, 0.0, axis=2)                        # :816-820
    conc = np.exp(np.einsum("nd,sdk->snk", m.x, beta) + alpha[:, None, :])   # :821-822
    pred = m.n_total[None, :, None] * conc / conc.sum(axis=2, keepdims=True)  # :824-827 DirichletMultinomial.mean
    conc_final = np.exp(m.x @ beta.mean(axis=0) + alpha.mean(axis=0))    # :836
    y_hat = conc_final / conc_final.sum(axis=1, keepdims=True) * m.n_total[:, None]  # :838
    return dict(sigma_d=sigma, b_offset=boff, ind_raw=iraw, alpha=alpha, ind=ind, b_raw=b_raw, beta=beta,
                concentration=conc, prediction=pred), y_hat


def inclusion_and_effects(beta: np.ndarray, est_fdr: float = 0.05):
    """result_classes.py:240-283: inclusion probability, non-zero mean, FDR threshold, final parameter.

    beta: [S, D, K] draws of ONE chain.
    """
    S, D, K = beta.shape
    inc = np.zeros(D * K)
    nzmean = np.zeros(D * K)
    for j in range(D):
        for i in range(K):
            b = beta[:, j, i]
            nz = np.abs(b) > 1e-3                                        # :249
            inc[j * K + i] = nz.sum() / S                                # :250
            nzmean[j * K + i] = b[nz].mean() if nz.any() else 0.0        # :252-255
    # direct posterior probability threshold (:262-274)
    incs = np.sort(inc[inc > 0])[::-1]
    threshold = 1.0
    for c in np.unique(incs):
        fdr = np.mean(1.0 - incs[incs >= c])
        if fdr < est_fdr:
            threshold = math.floor(c * 10 ** 3) / 10 ** 3
            break
    final = np.where(inc >= threshold, nzmean, 0.0)                      # :281-283
    return inc, nzmean, threshold, final


def hdi(v: np.ndarray, prob: float = 0.94):
    """arviz.hdi for a unimodal sample: shortest interval containing floor(prob*n) points."""
    v = np.sort(v[~np.isnan(v)])
    n = len(v)
    if n == 0:
        return math.nan, math.nan
    k = int(math.floor(prob * n))
    widths = v[k:] - v[:n - k]
    if len(widths) == 0:
        return math.nan, math.nan
    i = int(np.argmin(widths))
    return v[i], v[i + k]


def expected_samples(alpha_final: np.ndarray, beta_final: np.ndarray, y: np.ndarray):
    """result_classes.py:287-310, 337-340: expected sample and log2 fold change."""
    y_bar = np.mean(np.sum(y, axis=1))
    a = np.exp(alpha_final)
    alpha_sample = a / a.sum() * y_bar
    out, lfc = [], []
    for d in range(beta_final.shape[0]):
        b = np.exp(alpha_final + beta_final[d])
        b = b / b.sum() * y_bar
        out.append(b)
        lfc.append(np.log2(b / alpha_sample))
    return alpha_sample, np.array(out), np.array(lfc)
