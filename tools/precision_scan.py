"""fp32 value / gradient error of the device path vs the fp64 oracle as the concentrations grow (tools).
K1 kernel with one warp per chain (per-thread partial sums as in the samplers), element path and pair/item path;
the case/control kernel's traced log-density at its traced states."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import synthetic_dataset, random_states
from oracle import np_oracle as o
from sccoda_b200.engine import DeviceProblem

x, y = synthetic_dataset(7, N=20, K=20, D=1)
m = o.prepare(x, y, 19)
rs = np.random.RandomState(0)
C = 2048                                     # enough chains for the planner to take one warp per chain
base = random_states(rs, 1, 20, C, scale=0.3)
for grouped in (False, True):
    prob = DeviceProblem.from_arrays(x, y, 19, chains=C, dtype=np.float32, grouped=grouped)
    for shift in (0, 2, 4, 6, 8, 10, 12):
        th = base.copy(); th[:, -20:] += shift
        lp, g = prob.logp_grad(th[None]); lp = lp.cpu().numpy()[0]; g = g.double().cpu().numpy()[0]
        sel = slice(0, 16)
        lp_ref = np.array([o.logp(t, m) for t in th[sel]]); g_ref = np.array([o.grad(t, m) for t in th[sel]])
        cmax = np.exp(th[:, -20:]).max()
        print(f"grouped={int(grouped)} alpha shift {shift:2d}  c_max ~{cmax:9.3g}  |lp err| max {np.abs(lp[sel] - lp_ref).max():9.3g}   grad rel err "
              f"{(np.abs(g[sel] - g_ref) / np.abs(g_ref).max(1, keepdims=True)).max():9.3g}")
# case/control kernel: one HMC step from shifted states, traced logp vs oracle at the traced state
prob = DeviceProblem.from_arrays(x, y, 19, chains=64, dtype=np.float32)
for shift in (0, 4, 8, 12):
    th = base[:64].copy(); th[:, -20:] += shift
    out = prob.sample(th[None], prob.make_sampler(0, 2, 0, num_adapt=0, step_size=1e-4, seed=1))
    d = out.draws.double().cpu().numpy()[0]; lp = out.stat("target_log_prob").double().cpu().numpy()[0]
    ref = np.array([[o.logp(d[c, i], m) for i in range(2)] for c in range(16)])
    print(f"case/control kernel (kind {out.kernel_kind}) alpha shift {shift:2d}  |lp err| max {np.abs(lp[:16] - ref).max():9.3g}")
