"""CPU tests of the library's packer (`sccoda_pack`, csrc/pack.cpp) against the numpy statement of the same layout
(`_pack.pack_numpy`): integer arrays bit-exact, real arrays to the last bits (libm's lgamma against scipy's)."""
import ctypes
import dataclasses
import time

import numpy as np
import pytest

from conftest import haber_salm, synthetic_dataset
from sccoda_b200 import _native as nat
from sccoda_b200._pack import pack, pack_numpy


def _same(a, b):
    for f in dataclasses.fields(a):
        va, vb = getattr(a, f.name), getattr(b, f.name)
        if isinstance(va, np.ndarray):
            assert va.shape == vb.shape and va.dtype == vb.dtype, f.name
            if va.dtype.kind == "f":
                rtol = 2e-7 if va.dtype == np.float32 else 1e-13
                assert np.allclose(va, vb, rtol=rtol, atol=0.0, equal_nan=True), f.name
            else:
                assert np.array_equal(va, vb), f.name
        else:
            assert va == vb, f.name


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_haber_matches_numpy_packer(dtype):
    x, y, _ = haber_salm()
    x, y = x[[4, 0, 1, 5, 2, 3]], y[[4, 0, 1, 5, 2, 3]]
    _same(pack(x, y, 5, dtype=dtype), pack_numpy(x, y, 5, dtype=dtype))


@pytest.mark.parametrize("grouped", [None, True, False])
def test_ragged_batch_matches_numpy_packer(grouped):
    """pseudocounts, counts in {1..5}, non-integer counts, different references, unequal item counts per dataset"""
    xs, ys = zip(*[synthetic_dataset(s) for s in range(6)])
    ys = np.stack(ys)
    ys[1, :, 3] = 2
    ys[2, 3, 4] = 0
    ys[3, 2, 2] = 2.5
    ys[4, 0, :] = [1, 0, 2, 0.5, 1, 0, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16]
    refs = [19, 0, 3, 4, 5, 6]
    a, b = pack(np.stack(xs), ys, refs, grouped=grouped), pack_numpy(np.stack(xs), ys, refs, grouped=grouped)
    _same(a, b)
    assert a.NI.min() < a.NI.max()


def test_designs_and_layout_choice_match_numpy_packer():
    rs = np.random.RandomState(0)
    # continuous covariates with non-integer counts: one group per row, numpy's pairwise row sums reproduced bit for bit
    x = rs.randn(9, 2)
    y = rs.randint(1, 50, size=(9, 4)).astype(float) + rs.rand(9, 4)
    a, b = pack(x, y, 1, dtype=np.float64), pack_numpy(x, y, 1, dtype=np.float64)
    _same(a, b)
    assert np.array_equal(a.ntot, b.ntot)
    # four conditions (D = 3), a 2x2 factorial, K = 50 (two cell types per lane)
    _, y16 = synthetic_dataset(7, N=16, K=9)
    x4 = (np.arange(16)[:, None] % 4 == np.arange(1, 4)[None, :]).astype(float)
    _same(pack(x4, y16, 8), pack_numpy(x4, y16, 8))
    xf = np.stack([np.arange(16) % 2, (np.arange(16) // 2) % 2], axis=1).astype(float)
    _same(pack(xf, y16, 0, dtype=np.float64), pack_numpy(xf, y16, 0, dtype=np.float64))
    x50, y50 = synthetic_dataset(3, N=20, K=50)
    _same(pack(x50, y50, 49), pack_numpy(x50, y50, 49))
    # 40 rows with a continuous covariate: the automatic choice is the element layout in both
    x40, y40 = synthetic_dataset(0, N=40)
    xc = x40 + rs.randn(*x40.shape)
    a, b = pack(xc, y40, 19), pack_numpy(xc, y40, 19)
    assert a.grouped == b.grouped == 0
    _same(a, b)
    # no pseudocount: exact zeros stay and drop out of the pair/item layout
    y0 = y16.copy()
    y0[0, 0] = 0.0
    _same(pack(x4, y0, 8, dtype=np.float64, pseudocount=False), pack_numpy(x4, y0, 8, dtype=np.float64, pseudocount=False))


def test_oversized_item_tables_and_errors():
    rs = np.random.RandomState(1)
    N, K, D = 600, 80, 6
    x = rs.randn(N, D)
    y = rs.poisson(50, size=(N, K)).astype(float) + 1
    with pytest.raises(ValueError):
        pack(x, y, 3, grouped=True)
    a = pack(x, y, 3)
    assert a.grouped == 0 and a.Umax == N and int(a.NI.sum()) == 0
    _same(a, pack_numpy(x, y, 3))
    with pytest.raises(ValueError):
        pack(np.zeros((4, 1)), np.ones((5, 3)), 0)
    with pytest.raises(NameError):
        pack(np.zeros((5, 1)), np.ones((5, 3)), 3)
    # the C entry point reports the invalid reference itself (-52) when called directly
    h = ctypes.c_void_p()
    xx, yy, ref = np.zeros((1, 5, 1)), np.ones((1, 5, 3)), np.array([7], np.int32)
    rc = nat.lib().sccoda_pack(1, 5, 3, 1, xx.ctypes.data, yy.ctypes.data, ref.ctypes.data, nat.F32, 1, -1, 1, ctypes.byref(h))
    assert rc == -52 and not h.value and b"Reference index" in nat.lib().sccoda_last_error()


def test_thread_count_does_not_change_the_result_and_sweep_packs_fast():
    from sccoda_b200.util.data_generation import generate_sweep
    x, y, _ = generate_sweep(256, K=20, n_per_group=10, n_total=5000, seed0=1234)
    a = pack(x, y, 19, threads=1)
    t0 = time.perf_counter()
    b = pack(x, y, 19, threads=0)
    dt = time.perf_counter() - t0
    _same(a, b)
    for f in dataclasses.fields(a):
        if isinstance(getattr(a, f.name), np.ndarray):
            assert np.array_equal(getattr(a, f.name), getattr(b, f.name), equal_nan=True), f.name
    _same(a, pack_numpy(x, y, 19))
    assert dt < 1.0          # 256 datasets: milliseconds natively (the numpy packer needs ~0.13 s)
