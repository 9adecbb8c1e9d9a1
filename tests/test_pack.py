"""CPU tests of the host packer (the HBM data layout of include/sccoda_b200.h)."""
import math

import numpy as np
import pytest
from scipy.special import gammaln

from conftest import haber_salm, synthetic_dataset
from sccoda_b200._pack import pack
from oracle import np_oracle as o


def test_groups_sorted_and_counts_preserved():
    x, y, _ = haber_salm()
    x = x[[4, 0, 1, 5, 2, 3]]            # shuffle so that sorting does something
    y = y[[4, 0, 1, 5, 2, 3]]
    pk = pack(x, y, 5, dtype=np.float64)
    assert pk.Umax == 2 and pk.U[0] == 2
    assert np.all(np.diff(pk.grp[0]) >= 0)
    assert list(pk.rstart[0]) == [0, 4, 6]
    assert np.allclose(pk.xu[0, :, 0], [0.0, 1.0])
    assert np.allclose(pk.y[0], y[pk.row_order[0]])
    assert np.allclose(pk.ntot[0], y[pk.row_order[0]].sum(1))


def test_pseudocount_and_constant():
    x, y = synthetic_dataset(1, N=8, K=6)
    y[2, 3] = 0
    pk = pack(x, y, 0, dtype=np.float64)
    m = o.prepare(x, y, 0)
    assert 0.5 in pk.y[0]
    assert abs(pk.logp_const[0] - m.logp_const) < 1e-9
    D, K = 1, 6
    prior = D * math.log(2 / math.pi) - D * (K - 1) * math.log(2 * math.pi) - K * (math.log(5) + 0.5 * math.log(2 * math.pi))
    assert abs(pk.lconst[0] - (prior + m.logp_const)) < 1e-9


def test_fp32_constants():
    x, y = synthetic_dataset(2, N=6, K=5)
    y[0, 0] = 0
    y[1, 1] = 3
    pk = pack(x, y, 4, dtype=np.float32)
    ys = pk.ey[0].astype(np.float64)
    big = ys >= 6
    stir = (ys - 0.5) * np.log(ys) - ys + 0.5 * math.log(2 * math.pi)
    assert np.allclose(pk.eycst[0][big], (gammaln(ys) - stir)[big], atol=1e-7)
    assert np.allclose(pk.eycst[0][~big], gammaln(ys)[~big], rtol=1e-6)
    assert (~big).sum() >= 2


def test_element_list_layout():
    """Column-major, columns sorted by decreasing minimum count; packed index (u*K + k) | (n << 16)."""
    x, y = synthetic_dataset(4, N=8, K=6)
    y[3, 2] = 0          # column 2 becomes a small-count column
    y[:, 5] = 2          # column 5 too
    pk = pack(x, y, 0, dtype=np.float64)
    N, K = 8, 6
    ys, grp = pk.y[0], pk.grp[0]
    colmin = ys.min(axis=0)
    assert np.all(np.diff(colmin[pk.colperm[0]]) <= 0)
    assert pk.nbig[0] == N * (colmin >= 6).sum() == pk.nbig_min
    assert set(pk.colperm[0][pk.nbig[0] // N:]) >= {2, 5}
    for e in range(N * K):
        kk, n = divmod(e, N)
        k = pk.colperm[0][kk]
        assert pk.ey[0][e] == ys[n, k]
        assert pk.eidx[0][e] == (grp[n] * K + k) | (n << 16)
    assert np.all(pk.ey[0][:pk.nbig[0]] >= 6)


def test_continuous_covariates_have_one_group_per_row():
    rs = np.random.RandomState(0)
    x = rs.randn(9, 2)
    y = rs.randint(1, 50, size=(9, 4)).astype(float)
    pk = pack(x, y, 1)
    assert pk.Umax == 9 and np.array_equal(pk.grp[0], np.arange(9))
    assert np.allclose(pk.xu[0], x[pk.row_order[0]])


def test_batched_shapes_and_errors():
    xs, ys = zip(*[synthetic_dataset(s, N=10, K=5) for s in range(3)])
    pk = pack(np.stack(xs), np.stack(ys), [0, 1, 2])
    assert pk.y.shape == (3, 10, 5) and pk.P == 1 + 2 * 4 + 5 and list(pk.ref) == [0, 1, 2]
    with pytest.raises(ValueError):
        pack(np.zeros((4, 1)), np.ones((5, 3)), 0)
    with pytest.raises(NameError):
        pack(np.zeros((5, 1)), np.ones((5, 3)), 3)
