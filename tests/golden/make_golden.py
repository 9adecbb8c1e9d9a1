"""Generates tests/golden/k1_vectors.npz and tests/golden/sampler_vectors.npz.

The reference (theislab/scCODA) cannot be executed here: it needs tensorflow / tensorflow-probability, which are
not installable offline, and it holds no log-density / gradient / per-step sampler vectors of its
own.  These fixtures are therefore produced by the cited numpy restatement (oracle/np_oracle.py) and are
cross-checked AT GENERATION TIME against an independent evaluation of the same density written directly from
sccoda/model/scCODA_model.py:702-751 with torch float64 + autograd.  They pin three things:
  * the oracle itself (tests/test_golden.py, CPU): any later edit of oracle/ must reproduce them bit-closely;
  * the C restatement (same test);
  * the CUDA path (tests/test_gpu_golden.py, `-m gpu`): fp64 build to 1e-9, fp32 build within the north-star
    tolerance (1e-3 relative), element path and pair/item path, plus the first transitions of the three samplers.

Run from the repository root:  python tests/golden/make_golden.py
"""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import haber_salm, random_states, synthetic_dataset  # noqa: E402
from oracle import np_oracle as o  # noqa: E402


def torch_logp_grad(x, y, ref, theta):
    """Independent fp64 evaluation, written from scCODA_model.py:702-751 (distributions spelled out)."""
    import torch
    y = np.array(y, dtype=np.float64)
    if np.count_nonzero(y) != y.size:
        y[y == 0] = 0.5                                                       # :94-96
    xt, yt = torch.tensor(x, dtype=torch.float64), torch.tensor(y, dtype=torch.float64)
    n_total = yt.sum(1)                                                       # :99-100
    N, K = y.shape
    D = x.shape[1]
    th = torch.tensor(theta, dtype=torch.float64, requires_grad=True)
    sigma = th[:D].reshape(D, 1)
    boff = th[D:D + D * (K - 1)].reshape(D, K - 1)
    iraw = th[D + D * (K - 1):D + 2 * D * (K - 1)].reshape(D, K - 1)
    alpha = th[D + 2 * D * (K - 1):]
    lp = (math.log(2 / math.pi) - torch.log1p(sigma ** 2)).sum()              # HalfCauchy(0,1)   :703-707
    lp = lp + (-0.5 * boff ** 2 - 0.5 * math.log(2 * math.pi)).sum()          # Normal(0,1)       :709-714
    lp = lp + (-0.5 * iraw ** 2 - 0.5 * math.log(2 * math.pi)).sum()          # Normal(0,1)       :717-722
    ind = torch.sigmoid(50.0 * iraw)                                          # :724-725
    beta_t = ind * (sigma * boff)                                             # :727-729
    beta = torch.cat([beta_t[:, :ref], torch.zeros(D, 1, dtype=torch.float64), beta_t[:, ref:]], dim=1)   # :732-734
    lp = lp + (-0.5 * (alpha / 5.0) ** 2 - math.log(5.0) - 0.5 * math.log(2 * math.pi)).sum()            # :736-741
    conc = torch.exp(alpha + xt @ beta)                                       # :743
    S = conc.sum(1)
    lp = lp + (torch.lgamma(conc + yt) - torch.lgamma(conc)).sum() - (torch.lgamma(S + n_total) - torch.lgamma(S)).sum()
    lp = lp + (torch.lgamma(n_total + 1)).sum() - torch.lgamma(yt + 1).sum()  # DirichletMultinomial :746-751
    lp.backward()
    return float(lp.detach()), th.grad.numpy()


def k1_cases():
    x, y, _ = haber_salm()
    yield "haber_ref5", x, y, 5
    yield "haber_ref0", x, y, 0
    x, y = synthetic_dataset(7, N=20, K=20, D=1)
    y[0, 0] = 0.0
    yield "synth_20x20_d1", x, y, 19
    x, y = synthetic_dataset(5, N=12, K=9, D=3)
    y[3, 4] = 0.0
    yield "synth_12x9_d3", x, y, 2
    rs = np.random.RandomState(5)
    x, y = synthetic_dataset(11, N=16, K=12, D=2)
    yield "continuous_16x12_d2", x + 0.3 * rs.randn(*x.shape), y, 4
    x, y = synthetic_dataset(9, N=14, K=6, D=1)
    y[1, 1], y[2, 2], y[3, 3], y[4, 4], y[5, 0], y[6, 1], y[7, 2] = 1, 2, 3, 4, 5, 2.5, 0.25
    yield "small_counts_14x6", x, y, 5


def main():
    out = {}
    for name, x, y, ref in k1_cases():
        D, K = x.shape[1], y.shape[1]
        th = random_states(np.random.RandomState(len(name)), D, K, 6)
        th[0, :D] = 1.0
        th[0, D:] = 0.0
        th[1, -K:] += 5.0                       # large concentrations
        m = o.prepare(x, y, ref)
        lp = np.array([o.logp(t, m) for t in th])
        g = np.array([o.grad(t, m) for t in th])
        for i, t in enumerate(th):
            lp_t, g_t = torch_logp_grad(x, y, ref, t)
            assert abs(lp_t - lp[i]) <= 1e-10 * max(1.0, abs(lp_t)), (name, i, lp_t, lp[i])
            assert np.abs(g_t - g[i]).max() <= 1e-9 * max(1.0, np.abs(g_t).max()), (name, i)
        out[name + "/x"], out[name + "/y"], out[name + "/ref"] = x, y, np.int64(ref)
        out[name + "/theta"], out[name + "/logp"], out[name + "/grad"] = th, lp, g
    # SURVEY §8a anchor
    x, y, _ = haber_salm()
    m = o.prepare(x, y, 5)
    th0 = np.zeros(23)
    th0[0] = 1.0
    assert abs(o.logp(th0, m) - (-260.45211233545143)) < 1e-9
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "k1_vectors.npz"), **out)

    # first transitions of the three samplers on Haber/Salmonella (Philox streams of oracle/np_oracle.py)
    smp = {}
    rs = np.random.RandomState(0)
    init = np.concatenate([np.ones(1), rs.randn(7), np.zeros(7), rs.randn(8)])
    smp["init"] = init
    for method, steps in ((0, 30), (1, 8), (2, 6)):
        d, st = o.sample_chain(m, init, method, steps, 0, num_adapt=steps, seed=7, chain=3)
        smp[f"m{method}/draws"] = d
        for k, v in st.items():
            smp[f"m{method}/{k}"] = v
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "sampler_vectors.npz"), **smp)
    print("wrote", {k: v.shape for k, v in list(out.items())[:6]}, "...")


if __name__ == "__main__":
    main()
