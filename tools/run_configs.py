"""Run the five BASELINE.json configurations once through the public API / engine and report device times
(tools, not product; results are copied into profiles/)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sccoda_b200 import _native as nat
from sccoda_b200 import datasets, scheduler as sch
from sccoda_b200._pack import pack
from sccoda_b200.engine import DeviceProblem
from sccoda_b200.util import cell_composition_data as dat, comp_ana as mod, data_generation as gen

out = {}
quick = "--quick" in sys.argv

# ---- config 1: synthetic case-control 2x5, K=5, sample_hmc(20000, 5000), 1 chain
np.random.seed(1234)
b, w = gen.b_w_from_abs_change(np.array([200.] * 5), np.array([50, 0, 0, 0, 0]), 1000)
data = gen.generate_case_control(cases=1, K=5, n_total=1000, n_samples=[5, 5], b_true=b, w_true=w[None, :])
model = mod.CompositionalAnalysis(data, formula="x_0", reference_cell_type=4)
t0 = time.time(); res = model.sample_hmc(verbose=False, seed=1); wall = time.time() - t0
out["cfg1_case_control_hmc_1chain"] = dict(kernel_ms=model._last_run["kernel_ms"], wall_s=wall, acc=res.sampling_stats["acc_rate"],
                                          grad_evals_per_s=25000 * 10 / model._last_run["kernel_ms"] * 1e3,
                                          W=model._last_run["warps_per_chain"], effect_cell0=float(res.effect_df["Final Parameter"].iloc[0]),
                                          true_effect=float(w[0]))

# ---- config 2: Haber, all 10 samples, formula Condition (D=3), sample_hmc_da, 4 chains
df = datasets.haber()
d = dat.from_pandas(df, covariate_columns=["Mouse"])
d.obs["Condition"] = d.obs["Mouse"].str.replace(r"_[0-9]", "", regex=True)
model = mod.CompositionalAnalysis(d, formula="Condition", reference_cell_type="Goblet")
t0 = time.time(); res = model.sample_hmc_da(verbose=False, seed=2, num_chains=4); wall = time.time() - t0
out["cfg2_haber_hmc_da_4chains"] = dict(kernel_ms=model._last_run["kernel_ms"], wall_s=wall, acc=res.sampling_stats["acc_rate"],
                                       grad_evals_per_s=4 * 25000 * 10 / model._last_run["kernel_ms"] * 1e3, D=model.D,
                                       W=model._last_run["warps_per_chain"],
                                       credible=[str(i) for i, v in res.credible_effects().items() if v])

# ---- config 3: N=200, K=50, D=4, 64 chains, sample_nuts
rs = np.random.RandomState(1234)
N, K, D = 200, 50, 4
x = rs.randint(0, 2, size=(N, D)).astype(float)
bt = rs.uniform(-3, 3, size=K)
wt = np.zeros((D, K))
for dd in range(D):
    wt[dd, rs.choice(K - 1, 3, replace=False)] = rs.choice([0.3, 0.5, 1.0], 3)
logits = x @ wt + bt + np.sqrt(0.05) * rs.randn(N, K)
p = np.exp(logits - logits.max(1, keepdims=True)); p /= p.sum(1, keepdims=True)
y = np.stack([rs.multinomial(5000, p[n]) for n in range(N)]).astype(float)
pk = pack(x, y, K - 1, dtype=np.float32)
C = 64
nr, nb = (300, 200) if quick else (2000, 1000)
init = sch.initial_states(1, C, D, K, seed=3).astype(np.float32)
prob = DeviceProblem(pk, chains=C)
smp = prob.make_sampler(nat.METHOD_NUTS, nr, nb, max_tree_depth=10, seed=3)
o3 = prob.sample(init, smp)
lf = o3.stat("leapfrogs_taken").double()
ge = float(lf.sum().item()) * (nr + nb) / nr
out["cfg3_large_nuts_64chains"] = dict(kernel_ms=o3.kernel_ms, num_results=nr, num_burnin=nb, W=o3.warps_per_chain, grid=o3.grid,
                                      smem=o3.smem_bytes, U=int(pk.Umax), mean_leapfrogs=float(lf.mean().item()),
                                      grad_evals_per_s=ge / o3.kernel_ms * 1e3, finite=bool(torch.isfinite(o3.draws).all().item()),
                                      acc=float(torch.exp(torch.clamp(o3.stat("log_accept_ratio"), max=0)).mean().item()))

# ---- config 4: reference-cell-type sweep: K refits of one dataset in one launch
salm = df.iloc[[0, 1, 2, 3, 8, 9]]
ys = salm.iloc[:, 1:].to_numpy(float)
xs = np.array([0, 0, 0, 0, 1, 1.])[:, None]
Kh = 8
pk4 = pack(np.broadcast_to(xs, (Kh, 6, 1)), np.broadcast_to(ys, (Kh, 6, Kh)), np.arange(Kh), dtype=np.float32)
cfg = sch.SweepConfig(method=nat.METHOD_HMC, num_results=20000, num_burnin=5000, chains=4, seed=4)
sh = sch.sample_shard(None, None, None, cfg, 0, Kh, packed=pk4)
r4 = {k: v.cpu().numpy() for k, v in sch.shard_results(sh, cfg).items()}
thr, final = sch.credible_effect_calls(r4["inclusion_prob"], r4["mean_nonzero"])
out["cfg4_reference_sweep_8refits"] = dict(kernel_ms=sh.out.kernel_ms, grad_evals_per_s=Kh * 4 * 25000 * 10 / sh.out.kernel_ms * 1e3,
                                          enterocyte_credible_per_reference=[bool(final[r, 1] != 0) for r in range(Kh)],
                                          others_credible=int((final != 0).sum() - (final[:, 1] != 0).sum()))
print(json.dumps(out, indent=1))
