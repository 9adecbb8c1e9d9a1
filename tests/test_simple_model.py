"""SimpleModel (sccoda/model/other_models.py:31-282): the Dirichlet-multinomial model without spike-and-slab, run on
the scCODA kernels with sigma_d and ind_raw frozen.  CPU: the oracle's SimpleModel density against an independent
torch-autograd evaluation written from the reference; GPU: device trajectory, API and kernel variants."""
import math

import numpy as np
import pytest

from conftest import haber_salm, synthetic_dataset
from oracle import np_oracle as o


def _torch_simple(x, y, ref, b, alpha):
    """other_models.py:70-100 spelled out: b ~ N(0,1), alpha ~ N(0,5), DirMult(exp(alpha + x beta))."""
    import torch
    y = np.array(y, dtype=np.float64)
    if np.count_nonzero(y) != y.size:
        y[y == 0] = 0.5
    xt, yt = torch.tensor(x, dtype=torch.float64), torch.tensor(y, dtype=torch.float64)
    n = yt.sum(1)
    D = x.shape[1]
    bt = torch.tensor(b, dtype=torch.float64, requires_grad=True)
    at = torch.tensor(alpha, dtype=torch.float64, requires_grad=True)
    beta = torch.cat([bt[:, :ref], torch.zeros(D, 1, dtype=torch.float64), bt[:, ref:]], dim=1)
    conc = torch.exp(at + xt @ beta)
    S = conc.sum(1)
    lp = (-0.5 * bt ** 2 - 0.5 * math.log(2 * math.pi)).sum()
    lp = lp + (-0.5 * (at / 5.0) ** 2 - math.log(5.0) - 0.5 * math.log(2 * math.pi)).sum()
    lp = lp + (torch.lgamma(conc + yt) - torch.lgamma(conc)).sum() - (torch.lgamma(S + n) - torch.lgamma(S)).sum()
    lp = lp + torch.lgamma(n + 1).sum() - torch.lgamma(yt + 1).sum()
    lp.backward()
    return float(lp.detach()), bt.grad.numpy(), at.grad.numpy()


@pytest.mark.parametrize("shape", [(6, 8, 1), (12, 5, 3)])
def test_oracle_simple_model_density_and_gradient(shape):
    N, K, D = shape
    if N == 6:
        x, y, _ = haber_salm()
    else:
        x, y = synthetic_dataset(2, N=N, K=K, D=D)
        y[1, 2] = 0
    ref = 2
    m = o.prepare(x, y, ref)
    m.freeze = True
    rs = np.random.RandomState(0)
    for _ in range(3):
        b, alpha = rs.randn(D, K - 1) * 0.5, rs.randn(K)
        lp_t, gb_t, ga_t = _torch_simple(x, y, ref, b, alpha)
        th = o.simple_theta(b, alpha)
        assert abs(o.simple_logp(b, alpha, m) - lp_t) < 1e-9 * max(1.0, abs(lp_t))
        g = o.grad(th, m)
        DK1 = D * (K - 1)
        assert np.all(g[:D] == 0) and np.all(g[D + DK1:D + 2 * DK1] == 0)          # frozen parameters
        assert np.allclose(g[D:D + DK1].reshape(D, K - 1), gb_t, rtol=1e-9, atol=1e-9)
        assert np.allclose(g[-K:], ga_t, rtol=1e-9, atol=1e-9)


def test_frozen_oracle_chain_keeps_sigma_and_indicator():
    x, y, _ = haber_salm()
    m = o.prepare(x, y, 5)
    m.freeze = True
    rs = np.random.RandomState(1)
    th = o.simple_theta(rs.randn(1, 7), rs.randn(8))
    d, st = o.sample_chain(m, th, 0, 30, 0, num_adapt=30, seed=3, chain=0, target_accept=0.8)
    assert np.all(d[:, 0] == 1.0) and np.all(d[:, 8:15] == 1.0)
    assert st["is_accepted"].mean() > 0.5 and np.std(d[:, -8:], axis=0).min() > 0


@pytest.mark.gpu
@pytest.mark.parametrize("grouped", [False, True])
def test_device_frozen_trajectory_follows_oracle_fp64(grouped):
    from sccoda_b200.engine import DeviceProblem
    x, y, _ = haber_salm()
    m = o.prepare(x, y, 5)
    m.freeze = True
    rs = np.random.RandomState(1)
    th = o.simple_theta(rs.randn(1, 7), rs.randn(8))
    d_ref, s_ref = o.sample_chain(m, th, 0, 30, 0, num_adapt=30, seed=3, chain=2, target_accept=0.8)
    prob = DeviceProblem.from_arrays(x, y, 5, chains=1, dtype=np.float64, grouped=grouped)
    smp = prob.make_sampler(0, 30, 0, num_adapt=30, seed=3, chain_id0=2, target_accept=0.8, freeze_spike_slab=True)
    out = prob.sample(th[None, None], smp)
    d = out.draws.cpu().numpy()[0, 0]
    assert np.abs(d - d_ref).max() < 1e-9
    assert np.array_equal(out.stat("is_accepted").cpu().numpy()[0, 0], s_ref["is_accepted"])
    assert np.allclose(out.stat("target_log_prob").cpu().numpy()[0, 0], s_ref["target_log_prob"], rtol=1e-10)


@pytest.mark.gpu
def test_simple_model_api_on_haber():
    """SimpleModel.sample_hmc through the drop-in API: result contents of other_models.py:186-220, the Enterocyte
    effect of the Salmonella condition, target_log_prob == SimpleModel's own log-density at the traced states."""
    from sccoda_b200 import datasets
    from sccoda_b200.model.other_models import SimpleModel
    from sccoda_b200.util import cell_composition_data as dat
    df = datasets.haber().iloc[[0, 1, 2, 3, 8, 9]]
    data = dat.from_pandas(df, covariate_columns=["Mouse"])
    x = np.array([0, 0, 0, 0, 1, 1.0])[:, None]
    np.random.seed(3)
    model = SimpleModel(reference_cell_type=5, covariate_matrix=x, data_matrix=data.X, cell_types=list(data.var.index),
                        covariate_names=["Condition[T.Salm]"], formula="Condition")
    result = model.sample_hmc(num_results=4000, num_burnin=1000, seed=5)
    assert model._last_run["warps_per_chain"] == 1                              # the case/control kernel
    post = result.posterior
    assert set(post.data_vars) == {"b", "alpha", "beta", "concentration"}
    assert post["b"].shape == (1, 3000, 1, 7) and post["beta"].shape == (1, 3000, 1, 8)
    assert np.all(post["beta"][..., 5] == 0.0)
    assert 0.6 < result.sampling_stats["acc_rate"] < 0.98                        # adapted towards 0.8
    beta_mean = post["beta"][0].mean(axis=0)[0]
    assert beta_mean[1] > 0.8                                                    # Enterocyte (tutorial: ~1.35)
    # traced log-density == oracle SimpleModel density at the traced state (fp32 device arithmetic)
    m = o.prepare(x, np.asarray(data.X, dtype=np.float64), 5)
    lp = np.asarray(result.sample_stats["target_log_prob"])[0]
    for i in (0, 1500, 2999):
        ref = o.simple_logp(post["b"][0, i], post["alpha"][0, i], m)
        assert abs(lp[i] - ref) < 2e-2, (lp[i], ref)
    alpha_df, beta_df = result.summary_prepare()
    assert alpha_df.shape[0] == 8 and beta_df.shape[0] == 8


@pytest.mark.gpu
def test_frozen_case_control_kernels_match_generic_statistically():
    """fp32: the case/control HMC kernel, the generic kernel and the oracle agree on a SimpleModel posterior."""
    from sccoda_b200.engine import DeviceProblem
    x, y = synthetic_dataset(21, N=12, K=6)
    m = o.prepare(x, y, 5)
    m.freeze = True
    C, nr, nb = 32, 1500, 500
    rs = np.random.RandomState(2)
    init = np.stack([o.simple_theta(rs.randn(1, 5), rs.randn(6)) for _ in range(C)])
    means = []
    for generic in (True, False):
        prob = DeviceProblem.from_arrays(x, y, 5, chains=C, dtype=np.float32, grouped=True)
        smp = prob.make_sampler(0, nr, nb, seed=8, warps_per_chain=1, generic_only=generic, target_accept=0.8,
                                freeze_spike_slab=True)
        out = prob.sample(init[None], smp)
        assert out.kernel_kind == (1 if generic else 3)
        d = out.draws.double().cpu().numpy()[0]
        assert np.all(d[..., 0] == 1.0) and np.all(d[..., 6:11] == 1.0)
        means.append((d[:, nb // 2:].mean(1), out.stat("is_accepted").double().mean().item()))
    (m0, a0), (m1, a1) = means
    se = np.hypot(m0.std(0, ddof=1), m1.std(0, ddof=1)) / np.sqrt(C)
    assert np.all(np.abs(m0.mean(0) - m1.mean(0)) < 5 * se + 0.02)
    assert abs(a0 - a1) < 0.03 and 0.6 < a1 < 0.95
