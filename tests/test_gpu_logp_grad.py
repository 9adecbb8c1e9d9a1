"""K1 parity: log-density and gradient of the CUDA path (through the C ABI) vs the CPU oracle at identical states.

Tolerances from BASELINE.json north_star: 1e-5 relative in fp64, 1e-3 relative in fp32.
"""
import numpy as np
import pytest

from conftest import haber_salm, random_states, synthetic_dataset

pytestmark = pytest.mark.gpu

REL_F64 = 1e-5
REL_F32 = 1e-3


def _oracle(x, y, ref, thetas):
    from oracle import np_oracle as o
    m = o.prepare(x, y, ref)
    lp = np.array([o.logp(t, m) for t in thetas])
    g = np.array([o.grad(t, m) for t in thetas])
    return lp, g


def _check(x, y, ref, dtype, rel, n_states=16, seed=0, grouped=None):
    from sccoda_b200.engine import DeviceProblem
    D, K = x.shape[1], y.shape[1]
    rs = np.random.RandomState(seed)
    thetas = random_states(rs, D, K, n_states)
    thetas[0, :D] = 1.0
    thetas[0, D:] = 0.0                      # the reference's deterministic part of init (scCODA_model.py:763-768)
    lp_ref, g_ref = _oracle(x, y, ref, thetas)
    prob = DeviceProblem.from_arrays(x, y, ref, chains=n_states, dtype=dtype, grouped=grouped)
    lp, g = prob.logp_grad(thetas[None])
    lp = lp.cpu().numpy()[0]
    g = g.double().cpu().numpy()[0]
    # value: relative to |logp|
    assert np.all(np.abs(lp - lp_ref) <= rel * np.abs(lp_ref)), (lp, lp_ref)
    # gradient: relative to the gradient's scale at that state (components may vanish individually)
    scale = np.abs(g_ref).max(axis=1, keepdims=True)
    err = np.abs(g - g_ref) / scale
    assert err.max() <= rel, err.max()
    return np.abs(lp - lp_ref).max(), err.max()


@pytest.mark.parametrize("dtype,rel", [(np.float64, REL_F64), (np.float32, REL_F32)])
@pytest.mark.parametrize("ref", [0, 3, 5, 7])
@pytest.mark.parametrize("grouped", [False, True])
def test_haber(dtype, rel, ref, grouped):
    x, y, _ = haber_salm()
    _check(x, y, ref, dtype, rel, grouped=grouped)


def test_haber_known_answer_f64():
    """SURVEY §8a anchor: logp = -260.45211233545143 at sigma=1, rest 0, Haber Salm rows, ref=5."""
    from sccoda_b200.engine import DeviceProblem
    x, y, _ = haber_salm()
    D, K = 1, 8
    th = np.zeros((1, 1, D + 2 * D * (K - 1) + K))
    th[..., 0] = 1.0
    prob = DeviceProblem.from_arrays(x, y, 5, chains=1, dtype=np.float64)
    lp, _ = prob.logp_grad(th)
    assert abs(lp.item() - (-260.45211233545143)) < 1e-8


@pytest.mark.parametrize("dtype,rel", [(np.float64, REL_F64), (np.float32, REL_F32)])
@pytest.mark.parametrize("shape", [(10, 5, 1), (20, 20, 1), (12, 9, 3), (7, 2, 1), (40, 33, 2), (200, 50, 4)])
@pytest.mark.parametrize("grouped", [False, True])
def test_synthetic_shapes(dtype, rel, shape, grouped):
    N, K, D = shape
    x, y = synthetic_dataset(7, N=N, K=K, D=D)
    y[0, 0] = 0.0                              # exercise the pseudocount path (scCODA_model.py:94-96)
    _check(x, y, K - 1, dtype, rel, n_states=8, seed=3, grouped=grouped)


@pytest.mark.parametrize("dtype,rel", [(np.float64, REL_F64), (np.float32, REL_F32)])
@pytest.mark.parametrize("grouped", [False, True])
def test_small_and_non_integer_counts(dtype, rel, grouped):
    """Counts in {1..5}, the 0.5 pseudocount, non-integer and sub-unit values, a row with a tiny total."""
    x, y = synthetic_dataset(9, N=14, K=6, D=1)
    y[0, 0] = 0.0
    y[1, 1], y[2, 2], y[3, 3], y[4, 4], y[5, 0] = 1, 2, 3, 4, 5
    y[6, 1] = 2.5
    y[7, 2] = 0.25
    y[8, 3] = 5.75
    y[9] = [1, 0, 2, 0.5, 1, 0]                # row total 4.5 (+ pseudocounts): the total pair gets a small count too
    _check(x, y, 5, dtype, rel, n_states=8, seed=2, grouped=grouped)


@pytest.mark.parametrize("dtype,rel", [(np.float64, REL_F64), (np.float32, REL_F32)])
def test_no_pseudocount_zero_counts(dtype, rel):
    """pack(pseudocount=False): exact zeros contribute nothing to the likelihood terms of their pair."""
    from sccoda_b200.engine import DeviceProblem
    from sccoda_b200._pack import pack
    from oracle import np_oracle as o
    x, y = synthetic_dataset(4, N=10, K=5, D=1)
    y[0, 0] = 0.0
    y[3, 2] = 0.0
    rs = np.random.RandomState(0)
    th = random_states(rs, 1, 5, 4)
    out = []
    for grouped in (False, True):
        prob = DeviceProblem(pack(x, y, 4, dtype=dtype, pseudocount=False, grouped=grouped), chains=4)
        lp, g = prob.logp_grad(th[None])
        out.append((lp.cpu().numpy()[0], g.double().cpu().numpy()[0]))
    assert np.all(np.isfinite(out[0][1])) and np.all(np.isfinite(out[1][1]))
    scale = np.abs(out[0][1]).max(axis=1, keepdims=True)
    assert (np.abs(out[0][1] - out[1][1]) / scale).max() <= rel


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("grouped", [False, True])
def test_continuous_covariates(dtype, grouped):
    """No two rows share a covariate vector: U == N, the general path (and the grouped path with one row per group)."""
    rs = np.random.RandomState(5)
    x, y = synthetic_dataset(11, N=16, K=12, D=2)
    x = x + 0.3 * rs.randn(*x.shape)
    _check(x, y, 4, dtype, REL_F64 if dtype == np.float64 else REL_F32, n_states=8, grouped=grouped)


def test_negative_sigma_is_minus_inf():
    """HalfCauchy on the unconstrained sigma_d: logp = -inf, prior gradient 0."""
    from sccoda_b200.engine import DeviceProblem
    from oracle import np_oracle as o
    x, y, _ = haber_salm()
    rs = np.random.RandomState(1)
    th = random_states(rs, 1, 8, 4)
    th[:, 0] = -0.3
    prob = DeviceProblem.from_arrays(x, y, 5, chains=4, dtype=np.float64)
    lp, g = prob.logp_grad(th[None])
    assert np.all(np.isneginf(lp.cpu().numpy()))
    m = o.prepare(x, y, 5)
    g_ref = np.array([o.grad(t, m) for t in th])
    assert np.allclose(g.cpu().numpy()[0], g_ref, rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("grouped", [False, True])
def test_many_datasets_batched(grouped):
    """B datasets x C chains in one launch, each against its own oracle evaluation."""
    from sccoda_b200.engine import DeviceProblem
    from sccoda_b200._pack import pack
    B, C, N, K, D = 6, 3, 20, 20, 1
    xs, ys = zip(*[synthetic_dataset(100 + b, N=N, K=K, D=D) for b in range(B)])
    refs = np.arange(B) % K
    rs = np.random.RandomState(9)
    th = random_states(rs, D, K, B * C).reshape(B, C, -1)
    prob = DeviceProblem(pack(np.stack(xs), np.stack(ys), refs, dtype=np.float64, grouped=grouped), chains=C)
    lp, g = prob.logp_grad(th)
    lp, g = lp.cpu().numpy(), g.cpu().numpy()
    for b in range(B):
        lp_ref, g_ref = _oracle(xs[b], ys[b], int(refs[b]), th[b])
        assert np.allclose(lp[b], lp_ref, rtol=1e-10)
        assert np.allclose(g[b], g_ref, rtol=1e-8, atol=1e-8)
