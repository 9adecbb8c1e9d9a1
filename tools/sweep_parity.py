"""Full-length parity of the device sweep against the CPU oracle on generate_sweep data (tools, not product):
B datasets x 4 chains, sample_hmc(20000, 5000) on both sides with independent random streams; compares posterior
inclusion probabilities, credible-effect calls, acceptance rates and step sizes.  Writes one JSON object."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import c_oracle, np_oracle as o
from sccoda_b200 import scheduler as sch
from sccoda_b200.util.data_generation import generate_sweep

ap = argparse.ArgumentParser()
ap.add_argument("--datasets", type=int, default=128)
ap.add_argument("--chains", type=int, default=4)
ap.add_argument("--K", type=int, default=20)
ap.add_argument("--results", type=int, default=20000)
ap.add_argument("--burnin", type=int, default=5000)
a = ap.parse_args()
B, C, K, nr, nb = a.datasets, a.chains, a.K, a.results, a.burnin

x, y, w_true = generate_sweep(B, K=K, n_per_group=10)
cfg = sch.SweepConfig(method=0, num_results=nr, num_burnin=nb, chains=C, seed=11)
t0 = time.time()
res, shard = sch.run_sweep(x, y, K - 1, cfg)
t_dev = time.time() - t0
ms = [o.prepare(x[b], y[b], K - 1) for b in range(B)]
init = sch.initial_states(B, C, 1, K, seed=5)
t0 = time.time()
d, st_o = c_oracle.sample(np.stack([m.x for m in ms]), np.stack([m.y for m in ms]), np.stack([m.n_total for m in ms]),
                          [m.logp_const for m in ms], [m.ref for m in ms], init, 0, nr, nb, seed=99)
t_orc = time.time() - t0
used = d[:, :, nb:, :]
lp_o = st_o["target_log_prob"][:, :, nb:]
sane = lp_o.max(axis=2) < np.median(lp_o, axis=(1, 2))[:, None] + 500.0       # oracle chains outside the precision-loss mode
keep = sane.sum(1) >= 2
sig, boff, iraw = used[..., 0:1], used[..., 1:K], used[..., K:2 * K - 1]
with np.errstate(over="ignore"):
    beta = 1.0 / (1.0 + np.exp(-50.0 * iraw)) * sig * boff
nzm = np.abs(beta) > 1e-3
inc_chain = nzm.mean(2)
wgt = sane[:, :, None] / np.maximum(sane.sum(1), 1)[:, None, None]
inc_o = np.concatenate([(inc_chain * wgt).sum(1), np.zeros((B, 1))], axis=1)
nz = (np.where(nzm, beta, 0.0).sum(2) * sane[:, :, None]).sum(1) / np.maximum((nzm.sum(2) * sane[:, :, None]).sum(1), 1)
nz_o = np.concatenate([nz, np.zeros((B, 1))], axis=1)
inc_d = res["inclusion_prob"]
mean_c = (inc_chain * wgt).sum(1, keepdims=True)
var_c = (((inc_chain - mean_c) ** 2) * sane[:, :, None]).sum(1) / np.maximum(sane.sum(1)[:, None] - 1, 1)
se = float(np.sqrt(np.mean((var_c / np.maximum(sane.sum(1), 1)[:, None])[keep])))
diff = (inc_d - inc_o)[keep][:, :K - 1]
thr_o, final_o = sch.credible_effect_calls(inc_o, nz_o, 0.05)
call_d, call_o = (res["final_effect"] != 0)[keep], (final_o != 0)[keep]
kt = np.argmax(w_true, axis=1)[keep]
rows = np.arange(int(keep.sum()))
strong = w_true[keep][rows, kt] >= 1.0
out = {
    "workload": f"{B} datasets x {C} chains, K=N={K}, sample_hmc({nr},{nb}), L=10; device fp32 (kernel kind {shard.out.kernel_kind}) vs C oracle fp64, independent streams",
    "device_wall_s": t_dev, "device_kernel_ms": shard.out.kernel_ms, "oracle_wall_s": t_orc,
    "oracle_chains_in_precision_loss_mode": float(1.0 - sane.mean()), "datasets_compared": int(keep.sum()),
    "entries": int(diff.size), "typical_standard_error_of_an_entry": se,
    "inclusion_prob_rms_diff": float(np.sqrt(np.mean(diff ** 2))), "inclusion_prob_max_abs_diff": float(np.abs(diff).max()),
    "inclusion_prob_mean_diff": float(diff.mean()),
    "inclusion_prob_correlation": float(np.corrcoef(inc_d[keep][:, :K - 1].ravel(), inc_o[keep][:, :K - 1].ravel())[0, 1]),
    "credible_call_disagreement_rate": float((call_d != call_o).mean()),
    "credible_calls_device": int(call_d.sum()), "credible_calls_oracle": int(call_o.sum()),
    "true_effect_called_device": float(call_d[rows, kt].mean()), "true_effect_called_oracle": float(call_o[rows, kt].mean()),
    "true_effect_called_device_strong": float(call_d[rows, kt][strong].mean()) if strong.any() else None,
    "true_effect_called_oracle_strong": float(call_o[rows, kt][strong].mean()) if strong.any() else None,
    "true_effect_same_call": float((call_d[rows, kt] == call_o[rows, kt]).mean()),
    "false_positive_rate_device": float((call_d.sum() - call_d[rows, kt].sum()) / (call_d.size - rows.size)),
    "false_positive_rate_oracle": float((call_o.sum() - call_o[rows, kt].sum()) / (call_o.size - rows.size)),
    "acc_rate_device": float(np.mean(res["acc_rate"])) if "acc_rate" in res else None,
    "acc_rate_oracle_sane_chains": float(st_o["is_accepted"][sane].mean()),
    "final_step_size_oracle_sane_chains": float(st_o["step_size"][:, :, -1][sane].mean()),
}
try:
    out["acc_rate_device"] = float(shard.out.stat("is_accepted").float().mean().item())
    out["final_step_size_device"] = float(shard.out.stat("step_size")[:, :, -1].float().mean().item())
except Exception as exc:                                                       # noqa: BLE001
    out["device_stats_error"] = repr(exc)
print(json.dumps(out, indent=1))
