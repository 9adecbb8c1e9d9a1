"""CPU tests: device-side R-hat / ESS formulas against a plain loop transcription of the arviz/Stan definitions."""
import numpy as np
import pytest
import torch
from scipy import stats

from sccoda_b200 import diagnostics as dg


def _split(a):
    h = a.shape[1] // 2
    return np.concatenate([a[:, :h], a[:, a.shape[1] - h:]], axis=0)


def _zscale(a):
    r = stats.rankdata(a, method="average").reshape(a.shape)
    return stats.norm.ppf((r - 0.375) / (a.size + 0.25))


def _autocov(v):
    n = len(v)
    v = v - v.mean()
    return np.array([np.dot(v[:n - t], v[t:]) / n for t in range(n)])


def ess_loop(a):
    """arviz.stats.diagnostics._ess on already z-scaled split chains (loop form)."""
    C, S = a.shape
    acov = np.stack([_autocov(c) for c in a])
    chain_mean = a.mean(1)
    mean_var = acov[:, 0].mean() * S / (S - 1.0)
    var_plus = mean_var * (S - 1.0) / S
    if C > 1:
        var_plus += chain_mean.var(ddof=1)
    rho = np.zeros(S)
    even, rho[0] = 1.0, 1.0
    odd = 1.0 - (mean_var - acov[:, 1].mean()) / var_plus
    rho[1] = odd
    t = 1
    while t < S - 3 and (even + odd) > 0.0:
        even = 1.0 - (mean_var - acov[:, t + 1].mean()) / var_plus
        odd = 1.0 - (mean_var - acov[:, t + 2].mean()) / var_plus
        if even + odd >= 0:
            rho[t + 1], rho[t + 2] = even, odd
        t += 2
    max_t = t - 2
    if even > 0:
        rho[max_t + 1] = even
    t = 1
    while t <= max_t - 2:
        if rho[t + 1] + rho[t + 2] > rho[t - 1] + rho[t]:
            rho[t + 1] = (rho[t - 1] + rho[t]) / 2.0
            rho[t + 2] = rho[t + 1]
        t += 2
    ess = C * S
    tau = -1.0 + 2.0 * rho[:max_t + 1].sum() + rho[max_t + 1:max_t + 2].sum()
    tau = max(tau, 1 / np.log10(ess))
    return ess / tau


def rhat_loop(a):
    def plain(z):
        S = z.shape[1]
        return np.sqrt((S * z.mean(1).var(ddof=1) / z.var(1, ddof=1).mean() + S - 1) / S)
    s = _split(a)
    return max(plain(_zscale(s)), plain(_zscale(np.abs(s - np.median(s)))))


def ar1(rs, C, S, phi, shift=0.0):
    x = np.zeros((C, S))
    e = rs.randn(C, S)
    for t in range(1, S):
        x[:, t] = phi * x[:, t - 1] + e[:, t]
    return x + shift * np.arange(C)[:, None]


@pytest.mark.parametrize("phi", [0.0, 0.5, 0.9, -0.4])
def test_ess_bulk_matches_loop_definition(phi):
    rs = np.random.RandomState(int(abs(phi) * 10))
    a = ar1(rs, 4, 1200, phi)
    ref = ess_loop(_zscale(_split(a)))
    got = dg.ess_bulk(torch.from_numpy(a)).item()
    assert abs(got - ref) / ref < 0.03, (got, ref)


def test_ess_is_batched():
    rs = np.random.RandomState(3)
    a = np.stack([ar1(rs, 2, 600, p) for p in (0.1, 0.7, 0.95)])
    got = dg.ess_bulk(torch.from_numpy(a)).numpy()
    ref = np.array([ess_loop(_zscale(_split(x))) for x in a])
    assert np.all(np.abs(got - ref) / ref < 0.05)
    assert got[0] > got[1] > got[2]


def test_rhat_matches_loop_definition_and_flags_bad_mixing():
    rs = np.random.RandomState(5)
    good = ar1(rs, 4, 800, 0.3)
    bad = ar1(rs, 4, 800, 0.3, shift=1.5)
    r_good = dg.rhat(torch.from_numpy(good)).item()
    r_bad = dg.rhat(torch.from_numpy(bad)).item()
    assert abs(r_good - rhat_loop(good)) < 1e-6
    assert abs(r_bad - rhat_loop(bad)) < 1e-6
    assert r_good < 1.02 < 1.2 < r_bad


def test_rhat_from_moments():
    rs = np.random.RandomState(7)
    a = rs.randn(3, 4, 500, 6) + 0.1 * rs.randn(3, 4, 1, 6)       # [B,C,S,P]
    mean = a.mean(2)
    m2 = ((a - mean[:, :, None]) ** 2).sum(2)
    got = dg.rhat_from_moments(torch.from_numpy(mean), torch.from_numpy(m2), 500).numpy()
    W = a.var(2, ddof=1).mean(1)
    Bn = mean.var(1, ddof=1)
    assert np.allclose(got, np.sqrt((499 / 500 * W + Bn) / W))


def test_hdi_matches_arviz_definition_on_ragged_rows():
    """Device HDI (torch) == the numpy restatement of az.hdi with skipna, including NaN-masked and empty rows."""
    import torch
    from oracle import np_oracle as o
    from sccoda_b200 import diagnostics as dg
    rs = np.random.RandomState(0)
    x = rs.gamma(2.0, 1.0, size=(3, 5, 400))
    x[0, 1, rs.rand(400) < 0.6] = np.nan
    x[1, 2, :] = np.nan
    x[2, 0, 1:] = np.nan                              # a single valid draw: no window fits
    x[2, 3, :397] = np.nan                            # three valid draws
    got = dg.hdi(torch.from_numpy(x), 0.94).numpy()
    for i in range(3):
        for j in range(5):
            lo, hi = o.hdi(x[i, j], 0.94)
            if np.isnan(lo):
                assert np.isnan(got[i, j]).all()
            else:
                assert np.allclose(got[i, j], [lo, hi])
    got90 = dg.hdi(torch.from_numpy(x[0, 0]), 0.9).numpy()
    assert np.allclose(got90, o.hdi(x[0, 0], 0.9))
