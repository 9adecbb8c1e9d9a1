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
    pk = pack(x, y, 0, dtype=np.float64, grouped=False)
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
    # grouped layout: the gradient no longer walks the list, so every count >= 6 is listed first (stable partition)
    pg = pack(x, y, 0, dtype=np.float64, grouped=True)
    assert pg.nbig[0] == (ys >= 6).sum() > pk.nbig[0]
    assert np.all(pg.ey[0][:pg.nbig[0]] >= 6) and np.all(pg.ey[0][pg.nbig[0]:] < 6)
    key = lambda p: sorted(zip(p.eidx[0].tolist(), p.ey[0].tolist(), p.eycst[0].tolist()))
    assert key(pg) == key(pk)                                     # same multiset of (index, count, constant) records


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


def _direct_bracket(vals, c):
    from scipy.special import digamma
    return sum(digamma(c + v) - digamma(c) for v in vals)


def _layout_bracket(pk, b, p, c):
    """Br_p evaluated from the pair / item layout exactly as csrc/grouped.cuh does (scipy fp64)."""
    from scipy.special import digamma
    from sccoda_b200._pack import ITEM_R
    slots = np.zeros(pk.NP * pk.NF + 1)
    for i in range(pk.NI[b]):
        if pk.ipair[b, i] == p:
            assert pk.ipos[b, i] // pk.NF == p and slots[pk.ipos[b, i]] == 0.0
            slots[pk.ipos[b, i]] = sum(digamma(c + pk.iy[b, j, i]) - digamma(c + 6.0) for j in range(ITEM_R))
    for q in range(pk.NG[b]):
        if pk.opair[b, q] == p:
            o0, no = int(pk.odesc[b, q]) & 0xffff, int(pk.odesc[b, q]) >> 16
            assert pk.opos[b, q] // pk.NF == p and slots[pk.opos[b, q]] == 0.0 and no > 0
            slots[pk.opos[b, q]] = sum(digamma(c + pk.oy[b, t]) - digamma(c + 6.0) for t in range(o0, o0 + no))
    a = c * (c + 5.0)
    co = pk.pcoef[b, p]
    return (co[0] * c + co[1]) / a + (co[2] * c + co[3]) / (a + 4.0) + (co[4] * c + co[5]) / (a + 6.0) \
        + slots[p * pk.NF:(p + 1) * pk.NF].sum()


@pytest.mark.parametrize("shape", [(12, 7, 1), (6, 8, 1), (30, 5, 2)])
def test_pair_item_layout_reproduces_brackets(shape):
    """sum_n [psi(c + y_n) - psi(c)] of every pair (cell types and row totals) from items + rational + odd counts."""
    N, K, D = shape
    x, y = synthetic_dataset(3, N=N, K=K, D=D)
    y[0, 0] = 0                 # pseudocount 0.5 -> odd count
    y[1, 1] = 3
    y[2, 2] = 1
    y[3, 3] = 5
    y[4, 4] = 2.5               # non-integer data -> odd count
    y[5, 1] = 4
    pk = pack(x, y, K - 1, dtype=np.float64, grouped=True)
    assert pk.grouped == 1 and pk.NImax % 2 == 0 and pk.NImax >= pk.NI[0]
    assert pk.iy.shape == (1, 5, pk.NImax) and pk.ipair.dtype == np.uint16 and pk.odesc.dtype == np.uint32
    assert np.all(pk.iy >= 6.0) and pk.NF >= 2 and pk.NF & (pk.NF - 1) == 0
    n_odd = int((pk.odesc[0, :pk.NG[0]] >> 16).sum())
    assert 2.5 in pk.oy[0, :n_odd] and 0.5 in pk.oy[0, :n_odd]
    rs = np.random.RandomState(1)
    U, Umax = int(pk.U[0]), pk.Umax
    for u in range(U):
        rows = range(pk.rstart[0, u], pk.rstart[0, u + 1])
        for k in range(K + 1):
            p = u * K + k if k < K else Umax * K + u
            vals = [pk.y[0, n, k] if k < K else pk.ntot[0, n] for n in rows]
            for c in (np.exp(rs.randn()), 1e-3, 250.0):
                ref = _direct_bracket(vals, c)
                assert abs(_layout_bracket(pk, 0, p, c) - ref) <= 1e-12 * max(1.0, abs(ref))
            # the data counts of the pair's items (inreal per item, padding behind them) are its counts >= 6
            mine = [i for i in range(pk.NI[0]) if pk.ipair[0, i] == p]
            data = sorted(float(pk.iy[0, j, i]) for i in mine for j in range(int(pk.inreal[0, i])))
            assert data == sorted(v for v in vals if v >= 6)
            assert all(pk.iy[0, j, i] == 6.0 for i in mine for j in range(int(pk.inreal[0, i]), 5))


def test_pair_item_layout_auto_mode():
    """Few covariate groups -> grouped gradient path (also a continuous covariate on a few dozen rows: the
    lane-per-cell-type kernels take up to 32 groups); one row per group beyond that -> element path."""
    x, y = synthetic_dataset(0)
    assert pack(x, y, 19).grouped == 1
    rs = np.random.RandomState(0)
    assert pack(x + rs.randn(*x.shape), y, 19).grouped == 1
    x40, y40 = synthetic_dataset(0, N=40)
    assert pack(x40 + rs.randn(*x40.shape), y40, 19).grouped == 0
    assert pack(x40 + rs.randn(*x40.shape), y40, 19, grouped=True).grouped == 1
    # items of a batch are padded to a common even count; unused item slots are harmless (count 6, pair 0)
    xs, ys = zip(*[synthetic_dataset(s) for s in range(4)])
    ys = np.stack(ys)
    ys[1, :, 3] = 2
    pk = pack(np.stack(xs), ys, 19)
    assert pk.NI.min() < pk.NI.max() <= pk.NImax
    b = int(np.argmin(pk.NI))
    assert np.all(pk.iy[b, :, pk.NI[b]:] == 6.0) and np.all(pk.ipair[b, pk.NI[b]:] == 0)
    assert np.all(pk.ipos[b, pk.NI[b]:] == pk.NP * pk.NF)            # padding items write to the spare slot


def test_auto_layout_never_raises_for_oversized_item_tables():
    """A continuous design with many rows and cell types overflows the 16-bit indices of the pair/item layout: forcing
    that layout raises, the automatic choice silently takes the element layout."""
    rs = np.random.RandomState(1)
    N, K, D = 600, 80, 6
    x = rs.randn(N, D)
    y = rs.poisson(50, size=(N, K)).astype(float) + 1
    with pytest.raises(ValueError):
        pack(x, y, 3, grouped=True)
    pk = pack(x, y, 3)
    assert pk.grouped == 0 and pk.Umax == N and int(pk.NI.sum()) == 0
