"""CPU tests: the oracle itself (no GPU).

The reference's own tests hold no log-density / gradient vector and TensorFlow cannot run here, so
the restatement is pinned against (a) an independent torch-autograd float64 evaluation of the density written
directly from sccoda/model/scCODA_model.py:702-751, (b) the SURVEY §8a known-answer anchor, (c) the C port, and
(d) — in test_oracle_statistical.py — the reference's end-to-end assertions.
"""
import math

import numpy as np
import pytest
import torch

from conftest import haber_salm, random_states, synthetic_dataset
from oracle import c_oracle, np_oracle as o


def torch_logp(theta, m):
    """Straight transcription of the density factors with torch ops; autograd gives the gradient."""
    D, K = m.D, m.K
    th = torch.tensor(theta, dtype=torch.float64, requires_grad=True)
    sig = th[:D]
    boff = th[D:D + D * (K - 1)].reshape(D, K - 1)
    iraw = th[D + D * (K - 1):D + 2 * D * (K - 1)].reshape(D, K - 1)
    alpha = th[-K:]
    ind = torch.exp(50 * iraw) / (1 + torch.exp(50 * iraw))
    beta = ind * (sig[:, None] * boff)
    beta = torch.cat([beta[:, :m.ref], torch.zeros(D, 1, dtype=torch.float64), beta[:, m.ref:]], dim=1)
    conc = torch.exp(alpha + torch.tensor(m.x) @ beta)
    y, n = torch.tensor(m.y), torch.tensor(m.n_total)
    lp = torch.distributions.HalfCauchy(torch.ones(D, dtype=torch.float64)).log_prob(sig).sum()
    lp = lp + torch.distributions.Normal(0., 1.).log_prob(boff).sum()
    lp = lp + torch.distributions.Normal(0., 1.).log_prob(iraw).sum()
    lp = lp + torch.distributions.Normal(0., 5.).log_prob(alpha).sum()
    S = conc.sum(1)
    lp = lp + (torch.lgamma(conc + y) - torch.lgamma(conc)).sum() + (torch.lgamma(S) - torch.lgamma(S + n)).sum()
    lp = lp + (torch.lgamma(n + 1).sum() - torch.lgamma(y + 1).sum())
    lp.backward()
    return lp.item(), th.grad.numpy()


def test_known_answer_anchor():
    x, y, _ = haber_salm()
    m = o.prepare(x, y, 5)
    th = np.concatenate([np.ones(1), np.zeros(7), np.zeros(7), np.zeros(8)])
    assert abs(o.logp(th, m) - (-260.45211233545143)) < 1e-9
    assert abs(m.logp_const - 8853.951147691852) < 1e-8


@pytest.mark.parametrize("shape", [(6, 8, 1), (10, 5, 1), (20, 20, 1), (12, 9, 3)])
def test_logp_grad_vs_torch_autograd(shape):
    N, K, D = shape
    if N == 6:
        x, y, _ = haber_salm()
    else:
        x, y = synthetic_dataset(3, N=N, K=K, D=D)
    m = o.prepare(x, y, K // 2)
    rs = np.random.RandomState(0)
    for th in random_states(rs, D, K, 6):
        lp_t, g_t = torch_logp(th, m)
        lp, g = o.logp_and_grad(th, m)
        assert abs(lp - lp_t) < 1e-9 * max(1.0, abs(lp_t))
        assert np.allclose(g, g_t, rtol=1e-10, atol=1e-10)
        lp_c, g_c = c_oracle.logp_grad(m.x, m.y, m.n_total, m.logp_const, m.ref, th)
        assert abs(lp_c - lp) < 1e-9 * max(1.0, abs(lp))
        assert np.allclose(g_c, g, rtol=1e-11, atol=1e-11)


def test_pseudocount_before_row_sums():
    """scCODA_model.py:94-100 — zeros become 0.5 and n_total includes them."""
    y = np.array([[0., 10.], [3., 7.]])
    m = o.prepare(np.array([[0.], [1.]]), y, 1)
    assert m.y[0, 0] == 0.5 and m.n_total[0] == 10.5 and m.n_total[1] == 10.0
    m2 = o.prepare(np.array([[0.], [1.]]), np.array([[1., 10.], [3., 7.]]), 1)
    assert m2.y[0, 0] == 1.0


def test_dimension_mismatch_raises():
    with pytest.raises(ValueError):
        o.prepare(np.zeros((3, 1)), np.ones((4, 2)), 0)


def test_negative_sigma():
    x, y, _ = haber_salm()
    m = o.prepare(x, y, 5)
    th = random_states(np.random.RandomState(2), 1, 8, 1)[0]
    th[0] = -0.2
    assert o.logp(th, m) == -math.inf
    thp = th.copy(); thp[0] = 0.2
    # likelihood part of d/dsigma is odd in sigma through b_raw; the prior part vanishes for sigma < 0
    g_neg, g_pos = o.grad(th, m), o.grad(thp, m)
    assert np.isfinite(g_neg).all() and np.isfinite(g_pos).all()


def test_philox_known_answer():
    """Random123 known-answer vectors for philox4x32-10."""
    assert o.philox4x32(0, 0, 0, 0, 0, 0) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    assert o.philox4x32(0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff) == \
        (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
    assert o.philox4x32(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0) == \
        (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)


@pytest.mark.parametrize("method,steps", [(o.METHOD_HMC, 60), (o.METHOD_HMC_DA, 12), (o.METHOD_NUTS, 8)])
def test_c_port_follows_numpy_restatement(method, steps):
    """Same Philox streams => the two restatements walk the same trajectory until chaos amplifies rounding."""
    x, y, _ = haber_salm()
    m = o.prepare(x, y, 5)
    rs = np.random.RandomState(0)
    init = np.concatenate([np.ones(1), rs.randn(7), np.zeros(7), rs.randn(8)])
    d_np, s_np = o.sample_chain(m, init, method, steps, 0, num_adapt=steps, seed=7, chain=3)
    d_c, s_c = c_oracle.sample(m.x[None], m.y[None], m.n_total[None], [m.logp_const], [m.ref], init[None, None],
                               method, steps, 0, num_adapt=steps, seed=7, chain_id0=3)
    assert np.abs(d_np - d_c[0, 0]).max() < 1e-6
    assert np.array_equal(s_np["leapfrogs_taken"], s_c["leapfrogs_taken"][0, 0])
    assert np.array_equal(s_np["is_accepted"], s_c["is_accepted"][0, 0])


def test_dual_averaging_freezes_after_adaptation():
    da = o.DualAveraging(0.01, 5)
    for _ in range(5):
        da.update(math.log(0.9))
    eps, la = da.eps, da.log_avg
    assert abs(eps - math.exp(la)) < 1e-15         # at t == num_adapt the averaged iterate is used
    da.update(math.log(0.1))
    assert da.eps == eps and da.log_avg == la


def test_inclusion_threshold_logic():
    """result_classes.py:240-283 on a hand-made beta trace."""
    S = 1000
    beta = np.zeros((S, 1, 3))
    beta[:, 0, 0] = 0.5                      # always included
    beta[:400, 0, 1] = 0.2                   # included 40 %
    inc, nz, thr, final = o.inclusion_and_effects(beta, est_fdr=0.05)
    assert np.allclose(inc, [1.0, 0.4, 0.0])
    assert np.allclose(nz, [0.5, 0.2, 0.0])
    assert thr == 1.0                        # c=0.4 gives fdr 0.3; c=1.0 gives fdr 0 < 0.05 -> floor(1.0)
    assert np.allclose(final, [0.5, 0.0, 0.0])


def test_hdi_shortest_interval():
    v = np.concatenate([np.linspace(0, 1, 940), np.linspace(5, 100, 60)])
    lo, hi = o.hdi(v, 0.94)
    assert lo == 0.0 and hi <= 5.0
