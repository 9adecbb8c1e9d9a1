"""Value and gradient of the kernels the sweeps actually run — cc_grad / cc_value of the case/control kernels (kinds 3, 4)
and mg_grad / mg_value of the multi-group kernels (kinds 5, 6), reached through `sccoda_logp_grad_as` — against the CPU
oracle at identical states.  Tolerance from BASELINE.json north_star: 1e-3 relative in fp32 (these kernels are fp32 only).
Reference density: sccoda/model/scCODA_model.py:702-751."""
import numpy as np
import pytest

from conftest import haber_salm, random_states, synthetic_dataset

pytestmark = pytest.mark.gpu

REL = 1e-3
KIND_CC, KIND_MG = 3, 5


def _check(x, y, ref, kind, n_states=12, seed=0, scale=0.5):
    from oracle import np_oracle as o
    from sccoda_b200 import _native as nat
    from sccoda_b200._pack import pack
    from sccoda_b200.engine import DeviceProblem
    D, K = x.shape[1], y.shape[1]
    rs = np.random.RandomState(seed)
    thetas = random_states(rs, D, K, n_states, scale=scale)
    thetas[0, :D] = 1.0
    thetas[0, D:] = 0.0                      # deterministic part of the reference's init (scCODA_model.py:763-768)
    thetas[1, 0] = -0.3                      # sigma_d < 0: -inf value, zero prior gradient
    m = o.prepare(x, y, ref)
    lp_ref = np.array([o.logp(t, m) for t in thetas])
    g_ref = np.array([o.grad(t, m) for t in thetas])
    pk = pack(x, y, ref, dtype=np.float32)
    assert pk.grouped == 1
    prob = DeviceProblem(pk, chains=n_states)
    lp, g = prob.logp_grad(thetas[None], kernel_kind=kind)
    lp, g = lp.cpu().numpy()[0], g.double().cpu().numpy()[0]
    fin = np.isfinite(lp_ref)
    assert np.array_equal(np.isfinite(lp), fin) and np.all(lp[~fin] == lp_ref[~fin])
    assert np.all(np.abs(lp[fin] - lp_ref[fin]) <= REL * np.abs(lp_ref[fin])), (lp, lp_ref)
    gscale = np.abs(g_ref).max(axis=1, keepdims=True)
    err = np.abs(g - g_ref) / gscale
    assert err.max() <= REL, err.max()
    # and the generic kernel agrees with the family's kernel to the same tolerance (both are fp32)
    lp1, g1 = prob.logp_grad(thetas[None], kernel_kind=1)
    assert (np.abs(g1.double().cpu().numpy()[0] - g) / gscale).max() <= REL
    return float(np.abs(lp[fin] - lp_ref[fin]).max()), float(err.max())


@pytest.mark.parametrize("kind", [KIND_CC, KIND_MG])
@pytest.mark.parametrize("ref", [0, 3, 5, 7])
def test_haber(kind, ref):
    x, y, _ = haber_salm()
    _check(x, y, ref, kind)


@pytest.mark.parametrize("kind", [KIND_CC, KIND_MG])
def test_baseline_config_1_and_5_datasets(kind):
    """config 1: case-control 2 x 5 samples, K = 5; config 5: datasets of the 1024-dataset sweep (K = N = 20), the
    generator the bench uses."""
    from sccoda_b200.util.data_generation import generate_sweep
    x, y = synthetic_dataset(1, N=10, K=5)
    _check(x, y, 4, kind)
    xs, ys, _ = generate_sweep(6, K=20, n_per_group=10, n_total=5000, seed0=1234)
    for b in range(6):
        _check(xs[b], ys[b], 19, kind, n_states=8, seed=b)


@pytest.mark.parametrize("kind", [KIND_CC, KIND_MG])
def test_pseudocounts_small_and_non_integer_counts(kind):
    x, y = synthetic_dataset(9, N=14, K=6, D=1)
    y[0, 0] = 0.0
    y[1, 1], y[2, 2], y[3, 3], y[4, 4], y[5, 0] = 1, 2, 3, 4, 5
    y[6, 1] = 2.5
    y[7, 2] = 0.25
    y[8, 3] = 5.75
    y[9] = [1, 0, 2, 0.5, 1, 0]                # row total 4.5 (+ pseudocounts): the total pair gets a small count too
    _check(x, y, 5, kind, seed=2)


@pytest.mark.parametrize("kind", [KIND_CC, KIND_MG])
def test_many_odd_counts_per_pair(kind):
    """The odd counts of a pair ride in its item slots as y + 6 (cc_stage / mg_stage): pairs whose padding does not hold
    them all (7 zeros beside 3 large counts: a second item appears), several DISTINCT odd values in one pair (0.25, 0.5,
    2.5, 5.75 -> one rational correction each), a pair with odd counts only, and a column of zeros."""
    x, y = synthetic_dataset(21, N=20, K=9, D=1)
    y[:7, 0] = 0                                   # group 0: 7 pseudocounts + 3 large counts
    y[10:, 1] = [0, 0.25, 2.5, 5.75, 0, 0.25, 3, 1, 0, 7]          # group 1: mixed odd / small / large
    y[:10, 2] = 0.5                                # group 0: odd counts only (the data already hold 0.5)
    y[:, 3] = 0                                    # a cell type that was never seen
    y[3, 4] = 1e-3                                 # a tiny non-integer count
    _check(x, y, 8, kind, seed=5)
    _check(x, y, 0, kind, n_states=6, seed=6)      # ... with the sparse column as the reference


@pytest.mark.parametrize("K", [2, 3, 17, 31])
@pytest.mark.parametrize("kind", [KIND_CC, KIND_MG])
def test_cell_type_counts_one_per_lane(kind, K):
    x, y = synthetic_dataset(5, N=12, K=K)
    y[0, 0] = 0.0
    _check(x, y, K - 1, kind, n_states=8, seed=K)
    _check(x, y, 0, kind, n_states=4, seed=K + 1)


@pytest.mark.parametrize("K,D", [(32, 1), (50, 2), (62, 3)])
def test_two_cell_types_per_lane(K, D):
    x, y = synthetic_dataset(6, N=16, K=K, D=D)
    y[1, 40 % K] = 0.0
    _check(x, y, K - 1, KIND_MG, n_states=6, seed=K, scale=0.3)
    _check(x, y, 33 % K, KIND_MG, n_states=4, seed=K + 1, scale=0.3)


@pytest.mark.parametrize("U", [1, 2, 3, 4, 5, 8, 13, 16, 31, 32])
def test_covariate_group_counts(U):
    """U distinct covariate rows (one-hot conditions for U <= 5, a continuous covariate otherwise), 2 rows per group."""
    rs = np.random.RandomState(U)
    N, K = 2 * U, 9
    _, y = synthetic_dataset(U, N=N, K=K)
    if U == 1:
        x = np.zeros((N, 1))
    elif U <= 5:
        x = (np.arange(N)[:, None] % U == np.arange(1, U)[None, :]).astype(float)      # D = U - 1 <= 4
    else:
        x = np.repeat(rs.randn(U, 2), 2, axis=0)                                       # D = 2, U groups of 2 rows
    y[0, 1] = 0.0
    _check(x, y, K - 1, KIND_MG, n_states=6, seed=U, scale=0.3)


def test_four_covariates_factorial():
    N, K = 32, 12
    _, y = synthetic_dataset(3, N=N, K=K)
    x = ((np.arange(N)[:, None] >> np.arange(4)[None, :]) & 1).astype(float)           # 2^4 groups, D = 4
    _check(x, y, 3, KIND_MG, n_states=6, seed=4, scale=0.3)


@pytest.mark.parametrize("kind", [1, KIND_CC, KIND_MG])
def test_value_with_some_large_concentrations(kind):
    """States in which a few cell types have concentrations of 300 ... 60,000 while the others stay small: the value pass
    takes the joint (cancellation-free) form for the elements of those pairs only — two counts at a time with the form
    chosen per element (lgamma_mixed_big2).  |logp - oracle| stays at the fp32 level of the ordinary states."""
    from oracle import np_oracle as o
    from sccoda_b200._pack import pack
    from sccoda_b200.engine import DeviceProblem
    from sccoda_b200.util.data_generation import generate_sweep
    xs, ys, _ = generate_sweep(3, K=20, n_per_group=10, n_total=5000, seed0=1234)
    ys[1, 4, 7] = 0
    rs = np.random.RandomState(3)
    C = 10
    prob = DeviceProblem(pack(xs, ys, 19, dtype=np.float32), chains=C)
    th = np.stack([random_states(rs, 1, 20, C, scale=0.4) for _ in range(3)])
    for c in range(C):
        big = rs.choice(20, size=1 + c % 3, replace=False)
        th[:, c, -20 + big] = rs.uniform(5.8, 11.0, size=big.size)            # c = 330 ... 60,000
    lp, g = prob.logp_grad(th, kernel_kind=kind)
    lp, g = lp.cpu().numpy(), g.double().cpu().numpy()
    for b in range(3):
        m = o.prepare(xs[b], ys[b], 19)
        for c in range(C):
            ref = o.logp(th[b, c], m)
            assert abs(lp[b, c] - ref) <= 2e-6 * abs(ref) + 5e-3, (kind, b, c, lp[b, c], ref)
            gr = o.grad(th[b, c], m)
            assert np.abs(g[b, c] - gr).max() <= REL * np.abs(gr).max()


def test_wrong_shape_is_refused():
    from sccoda_b200._pack import pack
    from sccoda_b200.engine import DeviceProblem
    x, y = synthetic_dataset(6, N=16, K=9, D=2)
    prob = DeviceProblem(pack(x, y, 0, dtype=np.float32), chains=2)
    th = np.zeros((1, 2, prob.P))
    th[..., :2] = 1.0
    with pytest.raises(RuntimeError, match="does not have the shape"):
        prob.logp_grad(th, kernel_kind=KIND_CC)                    # D == 2: not a case/control design
    prob64 = DeviceProblem(pack(x, y, 0, dtype=np.float64), chains=2)
    with pytest.raises(RuntimeError, match="does not have the shape"):
        prob64.logp_grad(th, kernel_kind=KIND_MG)                  # the lane-per-cell-type kernels are fp32 only


def test_concentration_guard_is_a_parameter():
    """Default: a state with a concentration above 1e6 (fp32) has a NaN value; with the guard lifted the value is finite
    and close to the oracle's (alpha = 14.2 -> c = 1.5e6)."""
    from oracle import np_oracle as o
    from sccoda_b200._pack import pack
    from sccoda_b200.engine import DeviceProblem
    x, y, _ = haber_salm()
    prob = DeviceProblem(pack(x, y, 5, dtype=np.float32), chains=1)
    th = np.zeros((1, 1, prob.P))
    th[..., 0] = 1.0
    th[..., -8:] = 12.0
    th[..., -3] = 14.2
    ref = o.logp(th[0, 0], o.prepare(x, y, 5))
    for kind in (1, KIND_CC, KIND_MG):
        lp, _ = prob.logp_grad(th, kernel_kind=kind)
        assert np.isnan(lp.item())
        lp, _ = prob.logp_grad(th, kernel_kind=kind, reject_conc_above=-1.0)
        assert abs(lp.item() - ref) <= REL * abs(ref), (kind, lp.item(), ref)
        lp, _ = prob.logp_grad(th, kernel_kind=kind, reject_conc_above=1e5)
        assert np.isnan(lp.item())
