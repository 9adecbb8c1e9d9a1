"""Short run of the NUTS kernel at the shape of BASELINE config 3 (N=200, K=50, D=4, 64 chains) for timing / ncu
captures (tools, not product)."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sccoda_b200 import _native as nat
from sccoda_b200 import scheduler as sch
from sccoda_b200._pack import pack
from sccoda_b200.engine import DeviceProblem

ap = argparse.ArgumentParser()
ap.add_argument("--chains", type=int, default=64)
ap.add_argument("--results", type=int, default=40)
ap.add_argument("--burnin", type=int, default=20)
ap.add_argument("--W", type=int, default=0)
ap.add_argument("--grouped", type=int, default=-1)
ap.add_argument("--depth", type=int, default=10)
ap.add_argument("--method", type=int, default=2)
ap.add_argument("--D", type=int, default=4)
ap.add_argument("--N", type=int, default=200)
ap.add_argument("--K", type=int, default=50)
a = ap.parse_args()

rs = np.random.RandomState(1234)
N, K, D = a.N, a.K, a.D
x = rs.randint(0, 2, size=(N, D)).astype(float)
bt = rs.uniform(-3, 3, size=K)
wt = np.zeros((D, K))
for dd in range(D):
    wt[dd, rs.choice(K - 1, 3, replace=False)] = rs.choice([0.3, 0.5, 1.0], 3)
logits = x @ wt + bt + np.sqrt(0.05) * rs.randn(N, K)
p = np.exp(logits - logits.max(1, keepdims=True)); p /= p.sum(1, keepdims=True)
y = np.stack([rs.multinomial(5000, p[n]) for n in range(N)]).astype(float)
pk = pack(x, y, K - 1, dtype=np.float32, grouped=None if a.grouped < 0 else bool(a.grouped))
init = sch.initial_states(1, a.chains, D, K, seed=3).astype(np.float32)
prob = DeviceProblem(pk, chains=a.chains)
smp = prob.make_sampler(a.method, a.results, a.burnin, max_tree_depth=a.depth, seed=3, warps_per_chain=a.W)
o3 = prob.sample(init, smp)
lf = o3.stat("leapfrogs_taken").double() if a.method == 2 else (o3.stat("leapfrogs_taken").double() * 0 + 10)
ge = float(lf.sum().item()) * (a.results + a.burnin) / a.results
print(f"grouped={pk.grouped} kind={o3.kernel_kind} W={o3.warps_per_chain} grid={o3.grid} smem={o3.smem_bytes} ms={o3.kernel_ms:.1f} "
      f"mean_leapfrogs={lf.mean().item():.1f} grad-evals/s={ge / o3.kernel_ms * 1e3:.4e}")
