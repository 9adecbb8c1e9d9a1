"""fp32 value / gradient error of the device path vs the fp64 oracle as the concentrations grow (tools)."""
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
base = random_states(rs, 1, 20, 8, scale=0.3)
prob = DeviceProblem.from_arrays(x, y, 19, chains=8, dtype=np.float32)
for shift in (0, 2, 4, 6, 8, 10, 12):
    th = base.copy(); th[:, -20:] += shift
    lp_ref = np.array([o.logp(t, m) for t in th]); g_ref = np.array([o.grad(t, m) for t in th])
    lp, g = prob.logp_grad(th[None]); lp = lp.cpu().numpy()[0]; g = g.double().cpu().numpy()[0]
    cmax = np.exp(th[:, -20:]).max()
    print(f"alpha shift {shift:2d}  c_max ~{cmax:9.3g}  |lp err| max {np.abs(lp - lp_ref).max():9.3g}   grad rel err "
          f"{(np.abs(g - g_ref) / np.abs(g_ref).max(1, keepdims=True)).max():9.3g}   grad abs err {np.abs(g - g_ref).max():9.3g}  |g|max {np.abs(g_ref).max():8.3g}")
