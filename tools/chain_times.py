"""Distribution of the per-chain run times inside one sampler launch (carry slots P+5 = ns, P+6 = SM id; tools)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sccoda_b200 import scheduler as sch
from sccoda_b200._pack import pack
from sccoda_b200.engine import DeviceProblem
from sccoda_b200.util.data_generation import generate_sweep

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
x, y, _ = generate_sweep(B, K=20, n_per_group=10)
pk = pack(x, y, 19, dtype=np.float32)
prob = DeviceProblem(pk, chains=4)
init = sch.initial_states(B, 4, 1, 20, seed=1).astype(np.float32)
smp = prob.make_sampler(0, 2000, 500, step_size=0.02, seed=3)
out = prob.sample(init, smp)
out = prob.sample(init, smp)
cr = out.carry.cpu().numpy()
P = pk.P
ns, sm = cr[..., P + 5] * 1e-6, cr[..., P + 6].astype(int)      # [B,C] ms
print(f"kernel {out.kernel_ms:.2f} ms; chain ms: min {ns.min():.2f} mean {ns.mean():.2f} median {np.median(ns):.2f} p90 {np.quantile(ns, .9):.2f} max {ns.max():.2f}")
per_sm = np.array([ns[sm == s].max() if (sm == s).any() else 0 for s in range(148)])
cnt = np.array([(sm == s).sum() for s in range(148)])
print("warps per SM: min", cnt.min(), "max", cnt.max(), " SM end time: min %.2f mean %.2f max %.2f" % (per_sm.min(), per_sm.mean(), per_sm.max()))
odd = np.array([(pk.odesc[b, :pk.NG[b]] >> 16).max() if pk.NG[b] else 0 for b in range(B)])
for v in np.unique(odd):
    print("datasets with max odd group", v, ":", (odd == v).sum(), "mean chain ms %.2f" % ns[odd == v].mean())
h, e = np.histogram(ns, bins=12)
print("histogram", list(zip(np.round(e[:-1], 1), h)))
# the slowest datasets: what sets them apart?
dr = out.draws.float()
amax = dr[..., -20:].amax(dim=(1, 2, 3)).cpu().numpy()                      # largest alpha_k seen by any chain of the dataset
sig = dr[..., 0]
beta_max = (sig.unsqueeze(-1) * dr[..., 1:20]).abs().amax(dim=(1, 2, 3)).cpu().numpy()
ds_ms = ns.max(axis=1)
order = np.argsort(-ds_ms)[:12]
zeros = (y == 0).sum(axis=(1, 2))
small = ((y > 0) & (y < 6)).sum(axis=(1, 2))
for b in order:
    print(f"dataset {b}: {ds_ms[b]:.2f} ms  max alpha {amax[b]:.2f} (c ~ {np.exp(amax[b]):.0f})  |sigma b|max {beta_max[b]:.2f}  zeros {zeros[b]} small {small[b]} odd-max {odd[b]}")
print("fast ones:")
for b in np.argsort(ds_ms)[:5]:
    print(f"dataset {b}: {ds_ms[b]:.2f} ms  max alpha {amax[b]:.2f} (c ~ {np.exp(amax[b]):.0f})  zeros {zeros[b]} small {small[b]} odd-max {odd[b]}")
