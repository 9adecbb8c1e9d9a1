import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def haber_salm():
    """Haber et al. Control(4) + Salmonella(2) subset used by the reference tests (tests/unit_tests.py:152-162)."""
    import pandas as pd
    df = pd.read_csv(os.path.join(ROOT, "sccoda_b200", "datasets", "haber_counts.csv"))
    sub = df.iloc[[0, 1, 2, 3, 8, 9]]
    y = sub.iloc[:, 1:].to_numpy(dtype=np.float64)
    x = np.array([0, 0, 0, 0, 1, 1], dtype=np.float64)[:, None]
    return x, y, list(df.columns[1:])


def synthetic_dataset(seed, N=20, K=20, D=1, n_total=5000, effect=1.0):
    """Small case-control dataset following the recipe of sccoda.util.data_generation.generate_case_control
    (MVN(x w + b, 0.05 I) -> softmax -> multinomial), with its own RandomState (no global RNG)."""
    rs = np.random.RandomState(seed)
    b = rs.uniform(-2, 2, size=K)
    w = np.zeros((D, K))
    w[0, seed % (K - 1)] = effect
    x = np.zeros((N, D))
    x[N // 2:, 0] = 1.0
    for d in range(1, D):
        x[:, d] = rs.randint(0, 2, size=N)
    y = np.zeros((N, K))
    for n in range(N):
        a = rs.multivariate_normal(x[n] @ w + b, 0.05 * np.eye(K))
        p = np.exp(a - a.max())
        p /= p.sum()
        y[n] = rs.multinomial(n_total, p)
    return x, y


def random_states(rs, D, K, n, scale=0.5):
    """Plausible parameter states: sigma_d > 0 mostly, small ind_raw so that sigmoid(50 ind_raw) is not saturated."""
    P = D + 2 * D * (K - 1) + K
    th = rs.randn(n, P) * scale
    th[:, :D] = np.abs(th[:, :D]) + 0.1
    th[:, D + D * (K - 1):D + 2 * D * (K - 1)] *= 0.05
    return th


@pytest.fixture(scope="session")
def haber():
    return haber_salm()
