"""Short single-launch run of the sampler kernel at the benchmark shape, for ncu captures (tools, not product)."""
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
from sccoda_b200.util.data_generation import generate_sweep

ap = argparse.ArgumentParser()
ap.add_argument("--datasets", type=int, default=1024)
ap.add_argument("--chains", type=int, default=4)
ap.add_argument("--K", type=int, default=20)
ap.add_argument("--n-per-group", type=int, default=10)
ap.add_argument("--results", type=int, default=100)
ap.add_argument("--burnin", type=int, default=20)
ap.add_argument("--method", type=int, default=0)
ap.add_argument("--W", type=int, default=0)
ap.add_argument("--cpc", type=int, default=0)
ap.add_argument("--reps", type=int, default=1)
ap.add_argument("--step-size", type=float, default=0.02)
ap.add_argument("--generic", type=int, default=0, help="1: never take the case/control kernel")
ap.add_argument("--grouped", type=int, default=-1, help="-1 auto, 0 element path, 1 pair/item path")
ap.add_argument("--conditions", type=int, default=2, help="> 2: one-hot coded conditions (D = conditions - 1), n-per-group rows each")
ap.add_argument("--factorial", type=int, default=0, help="F > 0: F binary covariates (D = F), 2^F groups of n-per-group rows")
a = ap.parse_args()

if a.factorial > 0:
    G = 1 << a.factorial
    x, y, _ = generate_sweep(a.datasets, K=a.K, n_per_group=G * a.n_per_group // 2)
    lv = np.arange(y.shape[1]) % G
    x = np.broadcast_to(((lv[:, None] >> np.arange(a.factorial)[None, :]) & 1).astype(float), (a.datasets, y.shape[1], a.factorial)).copy()
elif a.conditions > 2:
    x, y, _ = generate_sweep(a.datasets, K=a.K, n_per_group=a.conditions * a.n_per_group // 2)
    lv = np.arange(y.shape[1]) % a.conditions
    x = np.broadcast_to((lv[:, None] == np.arange(1, a.conditions)[None, :]).astype(float), (a.datasets, y.shape[1], a.conditions - 1)).copy()
else:
    x, y, _ = generate_sweep(a.datasets, K=a.K, n_per_group=a.n_per_group)
pk = pack(x, y, a.K - 1, dtype=np.float32, grouped=None if a.grouped < 0 else bool(a.grouped))
init = sch.initial_states(a.datasets, a.chains, x.shape[2], a.K, seed=1).astype(np.float32)
prob = DeviceProblem(pk, chains=a.chains)
smp = prob.make_sampler(a.method, a.results, a.burnin, step_size=a.step_size, seed=3, warps_per_chain=a.W, chains_per_cta=a.cpc, generic_only=bool(a.generic))
for _ in range(a.reps):
    out = prob.sample(init, smp)
ge = a.datasets * a.chains * (a.results + a.burnin) * 10
print(f"grouped={pk.grouped} kind={out.kernel_kind} W={out.warps_per_chain} cpc={out.chains_per_cta} grid={out.grid} smem={out.smem_bytes} ms={out.kernel_ms:.3f} "
      f"grad-evals/s={ge / out.kernel_ms * 1e3:.4e} acc={out.stat('is_accepted').mean().item():.3f}")
