"""Generates tests/golden/trapped_state.npz: the final state of an ORACLE chain (fp64 restatement of the reference's
arithmetic, no concentration guard) that ended in the precision-loss mode of lgamma(c+y) - lgamma(c)
(scCODA_model.py:746-751 evaluated in float64), for tests/test_precision_loss.py.

    python tests/golden/make_trapped_state.py

12 datasets of the benchmark sweep x 4 chains, sample_hmc(20000, 5000), C oracle, seed 99 (the run VERDICT r1 quotes:
19 of 48 chains end at log-densities of +2.7e5 ... +4.7e5 against -1.9e3 at the posterior)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import c_oracle, np_oracle as o                      # noqa: E402
from sccoda_b200 import scheduler as sch                         # noqa: E402
from sccoda_b200.util.data_generation import generate_sweep      # noqa: E402

B, C, K = 12, 4, 20
x, y, _ = generate_sweep(B, K=K, n_per_group=10, n_total=5000, seed0=1234)
ms = [o.prepare(x[b], y[b], K - 1) for b in range(B)]
init = sch.initial_states(B, C, 1, K, seed=99)
d, st = c_oracle.sample(np.stack([m.x for m in ms]), np.stack([m.y for m in ms]), np.stack([m.n_total for m in ms]),
                        [m.logp_const for m in ms], [K - 1] * B, init, 0, 20000, 5000, seed=99)
lp = st["target_log_prob"]
trapped = lp[:, :, -1] > 1e4
print("trapped chains:", int(trapped.sum()), "of", B * C)
b, c = [int(v) for v in np.argwhere(trapped)[0]]
# the traced step at which the chain entered the mode, the last state before it and the first state inside
t_in = int(np.argmax(lp[b, c] > 1e4))
# largest concentration of every retained draw of the chains that stayed out of the mode (and of the trapped chains
# BEFORE they entered it): where does the posterior live relative to the guards 1e6 (fp32) / 1e15 (fp64)?
K1 = K - 1
maxc = []
for bb in range(B):
    for cc in range(C):
        stop = int(np.argmax(lp[bb, cc] > 1e4)) if trapped[bb, cc] else lp.shape[2]
        th = d[bb, cc, 5000:stop]                                 # second burn-in slice, up to the entry into the mode
        if th.shape[0] == 0:
            continue
        with np.errstate(over="ignore"):
            beta = 1.0 / (1.0 + np.exp(-50.0 * th[:, 1 + K1:1 + 2 * K1])) * th[:, 0:1] * th[:, 1:1 + K1]
        eta = th[:, -K:].copy()
        eta[:, :K1] += np.maximum(beta, 0.0)                      # x in {0, 1}: the larger of the two groups
        maxc.append(np.exp(eta.max(axis=1)))
maxc = np.concatenate(maxc)
print("draws:", maxc.size, "max concentration:", maxc.max(), "share above 1e6:", float((maxc > 1e6).mean()))
out = dict(dataset=b, chain=c, n_draws_outside_mode=int(maxc.size), max_conc_outside_mode=float(maxc.max()),
           share_above_1e6=float((maxc > 1e6).mean()), max_conc_quantiles=np.quantile(maxc, [0.5, 0.99, 0.9999, 1.0]), theta=d[b, c, -1], logp_oracle=lp[b, c, -1], n_trapped=int(trapped.sum()), n_chains=B * C,
           step_in=t_in, theta_before=d[b, c, max(t_in - 1, 0)], theta_first=d[b, c, t_in],
           logp_before=lp[b, c, max(t_in - 1, 0)], logp_first=lp[b, c, t_in],
           posterior_logp_median=float(np.median(lp[~trapped])))
np.savez(os.path.join(os.path.dirname(os.path.abspath(__file__)), "trapped_state.npz"), **out)
print({k: (v if np.ndim(v) == 0 else np.shape(v)) for k, v in out.items()})
