"""Pins the claim behind the concentration guard (sccoda_sampler::reject_conc_above).

The reference evaluates the Dirichlet-multinomial likelihood as float64 differences of lgamma values
(sccoda/model/scCODA_model.py:746-751 -> [TFP] lbeta(c + y) - lbeta(c)).  A divergent HMC trajectory can propose a state
with concentrations of 1e200; there c + y == c in float64, the differences vanish and the "log-density" reads hundreds of
thousands of nats ABOVE the posterior although its true value is tens of millions of nats BELOW it — the proposal is
accepted and the chain never returns.  tests/golden/trapped_state.npz (tests/golden/make_trapped_state.py) holds such a
transition of the oracle (17 of 48 chains of that run end this way); mpmath supplies the true value."""
import os

import mpmath as mp
import numpy as np
from scipy.special import gammaln

from oracle import np_oracle as o
from sccoda_b200.util.data_generation import generate_sweep

HERE = os.path.dirname(os.path.abspath(__file__))


def _load():
    z = np.load(os.path.join(HERE, "golden", "trapped_state.npz"))
    K = 20
    x, y, _ = generate_sweep(int(z["dataset"]) + 1, K=K, n_per_group=10, n_total=5000, seed0=1234)
    return z, o.prepare(x[int(z["dataset"])], y[int(z["dataset"])], K - 1)


def _concentrations(theta, m):
    sigma, boff, iraw, alpha = o.split(theta, m)
    _, _, beta = o._beta_full(sigma, boff, iraw, m)
    with np.errstate(over="ignore"):
        return np.exp(alpha[None, :] + m.x @ beta)


def _dm_true(conc, m, dps=700):
    """sum_n { sum_k [lgamma(c+y) - lgamma(c)] - lgamma(S+n) + lgamma(S) } with `dps` decimal digits."""
    with mp.workdps(dps):
        tot = mp.mpf(0)
        for n in range(m.N):
            cs = [mp.mpf(float(v)) for v in conc[n]]
            S = mp.fsum(cs)
            for k in range(m.K):
                tot += mp.loggamma(cs[k] + mp.mpf(float(m.y[n, k]))) - mp.loggamma(cs[k])
            tot += mp.loggamma(S) - mp.loggamma(S + mp.mpf(float(m.n_total[n])))
        return float(tot)


def _priors(theta, m):
    sigma, boff, iraw, alpha = o.split(theta, m)
    return float(np.sum(o.LOG_2_OVER_PI - np.log1p(sigma ** 2)) + np.sum(-0.5 * boff ** 2 - 0.5 * o.LOG_2PI)
                 + np.sum(-0.5 * iraw ** 2 - 0.5 * o.LOG_2PI)
                 + np.sum(-0.5 * (alpha / 5.0) ** 2 - np.log(5.0) - 0.5 * o.LOG_2PI))


def test_reference_arithmetic_is_exact_at_posterior_states():
    z, m = _load()
    th = z["theta_before"]
    conc = _concentrations(th, m)
    assert conc.max() < 1e3
    true = _priors(th, m) + _dm_true(conc, m, dps=60) + m.logp_const
    assert abs(o.logp(th, m) - true) < 1e-8 * abs(true)
    assert abs(o.logp(th, m) - float(z["logp_before"])) < 1e-6


def test_float64_lgamma_differences_lose_every_digit_at_the_trapped_state():
    z, m = _load()
    th = z["theta"]
    conc = _concentrations(th, m)
    assert conc.max() > 1e150 and th[0] > 10.0                       # sigma_d = 24, effects of several hundred
    posterior = float(z["posterior_logp_median"])                     # about -1.9e3
    lp64 = o.logp(th, m)                                              # the reference formula in float64
    assert lp64 > posterior + 1e5                                     # reads 3.7e5 nats ABOVE the posterior ...
    true = _priors(th, m) + _dm_true(conc, m) + m.logp_const
    assert true < posterior - 1e7                                     # ... its true value lies 2e7 nats BELOW it
    # the same in TFP's own association, lbeta(c + y) - lbeta(c) with lbeta(v) = sum lgamma(v) - lgamma(sum v)
    with np.errstate(all="ignore"):
        lbeta = lambda v: gammaln(v).sum(axis=1) - gammaln(v.sum(axis=1))
        tfp_form = float(np.sum(lbeta(conc + m.y) - lbeta(conc)))
    assert not np.isfinite(tfp_form) or abs(tfp_form - (true - _priors(th, m) - m.logp_const)) > 1e6
    # the transition that entered the mode was accepted only because of that: from a log-density of -1.97e3 the TRUE
    # log acceptance ratio of the proposal is below -1e7
    th1 = z["theta_first"]
    true1 = _priors(th1, m) + _dm_true(_concentrations(th1, m), m) + m.logp_const
    assert true1 - float(z["logp_before"]) < -1e7 < 1e5 < float(z["logp_first"]) - float(z["logp_before"])


def test_guard_rejects_the_state_and_cuts_no_posterior_mass():
    z, m = _load()
    import dataclasses
    for limit in (1e6, 1e15):
        mg = dataclasses.replace(m, reject_conc_above=limit)
        assert np.isnan(o.logp(z["theta"], mg))
        assert o.logp(z["theta_before"], mg) == o.logp(z["theta_before"], m)
    # 536,136 retained draws of the 48 oracle chains outside the mode: the largest concentration anywhere is ~330,
    # 3.5 decades below the fp32 guard and 12.5 below the fp64 guard — the guards cut no posterior mass on this data
    assert int(z["n_draws_outside_mode"]) > 500000
    assert float(z["max_conc_outside_mode"]) < 1e3 and float(z["share_above_1e6"]) == 0.0


def test_guarded_c_oracle_keeps_every_chain():
    """The chain of the fixture, run again with the guard: it stays at the posterior."""
    from oracle import c_oracle
    from sccoda_b200 import scheduler as sch
    z, m = _load()
    b, c = int(z["dataset"]), int(z["chain"])
    init = sch.initial_states(12, 4, 1, 20, seed=99)[b:b + 1]
    args = (m.x[None], m.y[None], m.n_total[None], [m.logp_const], [19], init, 0, 20000, 5000)
    _, st = c_oracle.sample(*args, seed=99, chain_id0=4 * b, reject_conc_above=1e6)
    assert st["target_log_prob"].max() < float(z["posterior_logp_median"]) + 500.0
    _, st = c_oracle.sample(*args, seed=99, chain_id0=4 * b)
    assert st["target_log_prob"][0, c, -1] > 1e5                      # without it: the fixture's trajectory, trapped
