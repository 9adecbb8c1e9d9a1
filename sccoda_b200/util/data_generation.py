"""
Synthetic compositional data (input generator for the benchmark configs and the parity tests).

API and distributional recipe of sccoda/util/data_generation.py:32-248 — logits ~ MVN(x w + b, sigma),
proportions = softmax(logits), counts ~ Multinomial(n_total, proportions) — drawn from numpy's legacy global
generator in the same order, so that the reference's golden vector (tests/unit_tests.py:44-76,
np.random.seed(1234)) is reproduced exactly.  `generate_sweep` is the batched generator used for the
1024-replicate benchmark (one independent RandomState per dataset, no global state).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np
import pandas as pd

from .cell_composition_data import CompositionData


def _softmax(v: np.ndarray) -> np.ndarray:
    e = np.exp(v - np.max(v))
    return e / e.sum()


def sparse_effect_matrix(D: int, K: int, n_d, n_k) -> np.ndarray:
    """Effect matrix [D,K], zero except on a random n_d x n_k sub-grid filled with values from {0.3, 0.5, 1}
    (data_generation.py:206-248)."""
    rows = np.random.choice(range(D), size=n_d, replace=False)
    cols = np.random.choice(range(K), size=n_k, replace=False)
    levels = (0.3, 0.5, 1)
    w = np.zeros((D, K))
    for r in rows:
        for c in cols:
            w[r, c] = levels[int(np.random.choice(3, 1)[0])]
    return w


def generate_case_control(cases: int = 1, K: int = 5, n_total: int = 1000, n_samples: List[int] = [5, 5],
                          sigma: Optional[np.ndarray] = None, b_true: Optional[np.ndarray] = None,
                          w_true: Optional[np.ndarray] = None) -> CompositionData:
    """Case-control data with `cases` binary covariates; sample block i carries the binary expansion of i
    (data_generation.py:32-121).  Note the reference's D = cases**2 for the default effect matrix (:69)."""
    D = cases ** 2
    if b_true is None:
        b_true = np.random.uniform(-3, 3, size=K).astype(np.float64)
    if w_true is None:
        n_d = np.random.choice(range(D), size=1)
        n_k = np.random.choice(range(K), size=1)
        w_true = sparse_effect_matrix(D, K, n_d, n_k)
    if sigma is None:
        sigma = 0.05 * np.identity(K)

    n_rows = int(sum(n_samples))
    x = np.zeros((n_rows, cases))
    y = np.zeros((n_rows, K))
    row = 0
    for combo in range(2 ** cases):
        bits = [int(ch) for ch in format(combo, "b").zfill(cases)]
        for _ in range(n_samples[combo]):
            x[row] = bits
            logits = np.random.multivariate_normal(mean=x[row] @ w_true + b_true, cov=sigma).astype(np.float64)
            y[row] = np.random.multinomial(n_total, _softmax(logits).astype(np.float64))
            row += 1

    obs = pd.DataFrame(x.astype(np.float64), columns=["x_" + str(i) for i in range(cases)])
    obs.index = obs.index.astype(str)
    return CompositionData(y.astype(np.float64), obs, uns={"b_true": b_true, "w_true": w_true})


def b_w_from_abs_change(counts_before: np.ndarray = np.array([200, 200, 200, 200, 200]),
                        abs_change: np.ndarray = np.array([50, 0, 0, 0, 0]),
                        n_total: int = 1000) -> Tuple[np.ndarray, np.ndarray]:
    """Intercepts b and effects w for an absolute change of the differentially abundant types, the others sharing
    the remaining cells equally; w is shifted so that the last type has effect 0 (data_generation.py:124-173)."""
    counts_before = np.asarray(counts_before, dtype=np.float64)
    K = counts_before.shape[0]
    b = np.log(counts_before / n_total)
    after = counts_before + abs_change
    # a scalar abs_change is added to every type but only type 0 counts as changed (reference behaviour, :157-158)
    changed = np.where(np.atleast_1d(np.asarray(abs_change) != 0))[0]
    rest = np.array([k for k in range(K) if k not in changed], dtype=int)
    after[rest] = (n_total - after[changed].sum()) / len(rest)
    w = np.log(after / n_total) - b
    return b, w - w[K - 1]


def counts_from_first(b_0: int = 200, n_total: int = 1000, K: int = 5) -> np.ndarray:
    """Count vector with first entry b_0 and the other K-1 entries equal, summing to n_total (:176-203)."""
    out = np.full(K, (n_total - b_0) / (K - 1), dtype=np.float64)
    out[0] = b_0
    return out


def generate_sweep(B: int, K: int = 20, n_per_group: int = 10, n_total: int = 5000, seed0: int = 1234,
                   dataset0: int = 0):
    """B independent case-control datasets (D=1, N=2*n_per_group) for the replicate benchmark (SURVEY §8d cfg 5):
    dataset i uses RandomState(seed0 + i), b ~ U(-3,3), one true effect from {0.3,0.5,1} on cell type i mod (K-1),
    logits ~ N(x w + b, 0.05 I), counts ~ Multinomial(n_total, softmax).  Returns x [B,N,1], y [B,N,K], w_true [B,K].
    """
    N = 2 * n_per_group
    x = np.zeros((B, N, 1))
    x[:, n_per_group:, 0] = 1.0
    y = np.zeros((B, N, K))
    w_all = np.zeros((B, K))
    levels = np.array([0.3, 0.5, 1.0])
    for i in range(B):
        g = dataset0 + i
        rs = np.random.RandomState(seed0 + g)
        b = rs.uniform(-3, 3, size=K)
        w = np.zeros(K)
        w[g % (K - 1)] = levels[rs.randint(3)]
        logits = b[None, :] + x[i] * w[None, :] + np.sqrt(0.05) * rs.randn(N, K)
        p = np.exp(logits - logits.max(axis=1, keepdims=True))
        p /= p.sum(axis=1, keepdims=True)
        for n in range(N):
            y[i, n] = rs.multinomial(n_total, p[n])
        w_all[i] = w
    return x, y, w_all
