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


def run_c_port_ess(K, n_per_group, num_results, num_burnin, L, n_threads):
    """One sample_hmc(num_results, num_burnin) chain per thread on datasets 0..n_threads-1 of the sweep (C + OpenMP
    port); bulk ESS (sccoda_b200.diagnostics, the arviz / Stan definition) of each chain after the reference's second
    burn-in over alpha, beta (non-reference) and target_log_prob.  Returns a dict for bench.py."""
    import torch
    from oracle import c_oracle, np_oracle as o
    from sccoda_b200 import diagnostics as dg        # post-processing library code (torch sort / FFT), not the sampler
    B, D = n_threads, 1
    xs, ys, nts, consts = [], [], [], []
    for i in range(B):
        x, y = _dataset(i, K, n_per_group)
        m = o.prepare(x, y, K - 1)
        xs.append(m.x); ys.append(m.y); nts.append(m.n_total); consts.append(m.logp_const)
    rs = np.random.RandomState(0)
    init = np.zeros((B, 1, D + 2 * D * (K - 1) + K))
    init[..., :D] = 1
    init[..., D:D + D * (K - 1)] = rs.randn(B, 1, D * (K - 1))
    init[..., -K:] = rs.randn(B, 1, K)
    t0 = time.perf_counter()
    d, st = c_oracle.sample(np.stack(xs), np.stack(ys), np.stack(nts), consts, [K - 1] * B, init, 0, num_results, num_burnin,
                            num_leapfrog=L, seed=1, n_threads=n_threads, reject_conc_above=1e15)
    wall = time.perf_counter() - t0
    skip = min(num_burnin, num_results)
    d = d[:, :, skip:, :]
    K1 = K - 1
    with np.errstate(over="ignore"):
        beta = 1.0 / (1.0 + np.exp(-50.0 * d[..., 1 + K1:1 + 2 * K1])) * d[..., 0:1] * d[..., 1:1 + K1]
    coords = np.concatenate([d[..., -K:], beta, st["target_log_prob"][:, :, skip:, None]], axis=-1)   # [B,1,S,coord]
    ess = dg.ess_bulk(torch.as_tensor(coords).permute(0, 3, 1, 2).contiguous()).numpy()               # [B,coord]
    return {"chains": B, "cores": n_threads, "wall_s": wall, "sample": f"{B} chains (one per thread, datasets 0..{B - 1}) x "
            f"sample_hmc({num_results},{num_burnin}) L={L}, C + OpenMP port of the oracle, fp64",
            "grad_evals_per_s": B * (num_results + num_burnin) * L / wall,
            "ess_min_coord_per_chain_median": float(np.median(np.nanmin(ess, axis=1))),
            "ess_per_s_min_coord": float(np.nanmin(ess, axis=1).sum() / wall),
            "ess_per_s_median_coord": float(np.nanmedian(ess, axis=1).sum() / wall)}
