"""
ORACLE — TEST / MEASUREMENT INFRASTRUCTURE ONLY.

CPU-baseline workers for bench.py (`cpu_baseline` leg and `--impl reference`): time the restated sampler
(oracle/np_oracle.py, numpy/scipy fp64; or oracle/dm_oracle.c) on the host cores.  Never used by the product.
"""
from __future__ import annotations

import time

import numpy as np


def warm(_):
    from oracle import np_oracle  # noqa: F401  (import cost outside the timed region)
    import scipy.special  # noqa: F401
    _dataset(0, 5, 2)
    return 0


def _dataset(i, K, n_per_group):
    from sccoda_b200.util.data_generation import generate_sweep   # input generator only (not the path)
    x, y, _ = generate_sweep(1, K=K, n_per_group=n_per_group, n_total=5000, seed0=1234, dataset0=i)
    return x[0], y[0]


def run_numpy_chain(task):
    """One sample_hmc-style chain of `steps` transitions on dataset i; returns (grad_evals, seconds)."""
    i, K, n_per_group, steps, L = task
    from oracle import np_oracle as o
    x, y = _dataset(i, K, n_per_group)
    m = o.prepare(x, y, K - 1)
    rs = np.random.RandomState(i)
    D = 1
    init = np.concatenate([np.ones(D), rs.randn(D * (K - 1)), np.zeros(D * (K - 1)), rs.randn(K)])
    t0 = time.perf_counter()
    o.sample_chain(m, init, o.METHOD_HMC, steps, 0, num_adapt=steps, num_leapfrog=L, seed=1, chain=i, fast_rng=True)
    return steps * L, time.perf_counter() - t0


def run_c_port(K, n_per_group, steps, L, n_threads):
    """C + OpenMP port on `n_threads` chains in parallel; returns (grad-evals/s, grad_evals, seconds)."""
    from oracle import c_oracle, np_oracle as o
    B = n_threads
    xs, ys, nts, consts = [], [], [], []
    for i in range(B):
        x, y = _dataset(i, K, n_per_group)
        m = o.prepare(x, y, K - 1)
        xs.append(m.x); ys.append(m.y); nts.append(m.n_total); consts.append(m.logp_const)
    rs = np.random.RandomState(0)
    D = 1
    init = np.zeros((B, 1, D + 2 * D * (K - 1) + K))
    init[..., :D] = 1
    init[..., D:D + D * (K - 1)] = rs.randn(B, 1, D * (K - 1))
    init[..., -K:] = rs.randn(B, 1, K)
    args = (np.stack(xs), np.stack(ys), np.stack(nts), consts, [K - 1] * B, init, 0, steps, 0)
    c_oracle.sample(*args[:6], 0, 10, 0, num_adapt=10, num_leapfrog=L, seed=1, n_threads=n_threads)   # warm
    t0 = time.perf_counter()
    c_oracle.sample(*args, num_adapt=steps, num_leapfrog=L, seed=1, n_threads=n_threads)
    wall = time.perf_counter() - t0
    ge = B * steps * L
    return ge / wall, ge, wall
