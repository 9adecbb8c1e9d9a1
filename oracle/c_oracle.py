"""
ORACLE — TEST INFRASTRUCTURE ONLY.  ctypes loader for oracle/dm_oracle.c (see its header).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
NSTAT = 8
STAT_NAMES = ("target_log_prob", "log_accept_ratio", "is_accepted", "step_size", "diverging",
              "leapfrogs_taken", "energy", "reach_max_depth")


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libdm_oracle.so")
    src = os.path.join(_HERE, "dm_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libdm_oracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        dp = ctypes.POINTER(ctypes.c_double)
        ip = ctypes.POINTER(ctypes.c_int)
        _LIB.oracle_logp_grad.argtypes = [ctypes.c_int] * 4 + [dp, dp, dp, ctypes.c_double, dp, dp, dp]
        _LIB.oracle_sample.argtypes = ([ctypes.c_int] * 5 + [ip, dp, dp, dp, dp, dp] + [ctypes.c_int] * 6
                                       + [ctypes.c_double, ctypes.c_double, ctypes.c_uint64, ctypes.c_uint32,
                                          dp, dp, ctypes.c_int])
        _LIB.oracle_max_threads.restype = ctypes.c_int
        _LIB.oracle_set_reject_conc_above.argtypes = [ctypes.c_double]
        _LIB.oracle_set_reject_conc_above.restype = None
    return _LIB


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def logp_grad(x, y, n_total, logp_const, ref, theta):
    x = np.ascontiguousarray(x, np.float64); y = np.ascontiguousarray(y, np.float64)
    n_total = np.ascontiguousarray(n_total, np.float64); theta = np.ascontiguousarray(theta, np.float64)
    N, K = y.shape
    D = x.shape[1]
    lp = np.zeros(1)
    g = np.zeros(theta.shape[0])
    lib().oracle_logp_grad(N, K, D, int(ref), _dp(x), _dp(y), _dp(n_total), float(logp_const), _dp(theta), _dp(lp), _dp(g))
    return float(lp[0]), g


def sample(x, y, n_total, logp_const, ref, init, method, num_results, num_burnin, num_adapt=None,
           num_leapfrog=10, max_tree_depth=10, step_size=0.01, target_accept=0.75, seed=0, chain_id0=0,
           n_threads=0, reject_conc_above=None):
    """x[B,N,D] y[B,N,K] n_total[B,N] logp_const[B] ref[B] init[B,C,P] -> draws[B,C,S,P], stats dict of [B,C,S].

    reject_conc_above: None = the reference's arithmetic (no guard); a limit = states with a concentration above it
    are rejected, as the device kernels do (sccoda_sampler::reject_conc_above)."""
    lib().oracle_set_reject_conc_above(float(reject_conc_above) if reject_conc_above else 0.0)
    x = np.ascontiguousarray(x, np.float64); y = np.ascontiguousarray(y, np.float64)
    n_total = np.ascontiguousarray(n_total, np.float64)
    logp_const = np.ascontiguousarray(logp_const, np.float64)
    ref = np.ascontiguousarray(ref, np.int32)
    init = np.ascontiguousarray(init, np.float64)
    B, N, K = y.shape
    D = x.shape[2]
    C, P = init.shape[1], init.shape[2]
    assert P == D + 2 * D * (K - 1) + K
    if num_adapt is None:
        num_adapt = int(0.8 * num_burnin)
    draws = np.zeros((B, C, num_results, P))
    stats = np.zeros((B, C, num_results, NSTAT))
    lib().oracle_sample(B, C, N, K, D, ref.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), _dp(x), _dp(y), _dp(n_total),
                        _dp(logp_const), _dp(init), int(method), int(num_results), int(num_burnin), int(num_adapt),
                        int(num_leapfrog), int(max_tree_depth), float(step_size), float(target_accept),
                        int(seed), int(chain_id0), _dp(draws), _dp(stats), int(n_threads))
    lib().oracle_set_reject_conc_above(0.0)
    return draws, {k: stats[..., i] for i, k in enumerate(STAT_NAMES)}


def max_threads() -> int:
    return int(lib().oracle_max_threads())
