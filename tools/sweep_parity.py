"""Full-length parity of the device sweep against the CPU oracle on generate_sweep data (tools, not product):
B datasets x 4 chains, sample_hmc(20000, 5000) on both sides with independent random streams; compares posterior
inclusion probabilities, credible-effect calls, intercepts (alpha) and effects (beta), acceptance rates and step sizes.
Both sides sample the same target: the oracle runs with the concentration guard of the fp32 kernels (reject_conc_above =
1e6; --oracle-guard 0 switches it off and the oracle's trapped chains are then counted and left out).  A third run — the
device in fp64 with the fp64 default guard 1e15 — measures how much posterior mass has a concentration above 1e6.
Writes one JSON object."""
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
ap.add_argument("--oracle-guard", type=float, default=1e6)
ap.add_argument("--fp64-datasets", type=int, default=32, help="datasets of the fp64 device run that measures the mass above 1e6")
ap.add_argument("--conditions", type=int, default=2, help="> 2: one-hot coded conditions (D = conditions - 1), 5 rows each, "
                                                             "one true effect per condition")
a = ap.parse_args()
B, C, K, nr, nb = a.datasets, a.chains, a.K, a.results, a.burnin

if a.conditions > 2:
    D, N = a.conditions - 1, 5 * a.conditions
    lv = np.arange(N) % a.conditions
    x1 = (lv[:, None] == np.arange(1, a.conditions)[None, :]).astype(float)
    x = np.broadcast_to(x1, (B, N, D)).copy()
    y = np.zeros((B, N, K))
    w_true = np.zeros((B, D, K))
    for i in range(B):
        rs = np.random.RandomState(4321 + i)
        b0 = rs.uniform(-3, 3, size=K)
        for d in range(D):
            w_true[i, d, rs.randint(K - 1)] = rs.choice([0.5, 1.0, 1.5])
        logits = b0[None, :] + x1 @ w_true[i] + np.sqrt(0.05) * rs.randn(N, K)
        pr = np.exp(logits - logits.max(axis=1, keepdims=True))
        pr /= pr.sum(axis=1, keepdims=True)
        for n in range(N):
            y[i, n] = rs.multinomial(5000, pr[n])
else:
    D = 1
    x, y, w1 = generate_sweep(B, K=K, n_per_group=10)
    w_true = w1[:, None, :]
cfg = sch.SweepConfig(method=0, num_results=nr, num_burnin=nb, chains=C, seed=11)
t0 = time.time()
res, shard = sch.run_sweep(x, y, K - 1, cfg)
t_dev = time.time() - t0
ms = [o.prepare(x[b], y[b], K - 1) for b in range(B)]
init = sch.initial_states(B, C, D, K, seed=5)
t0 = time.time()
d, st_o = c_oracle.sample(np.stack([m.x for m in ms]), np.stack([m.y for m in ms]), np.stack([m.n_total for m in ms]),
                          [m.logp_const for m in ms], [m.ref for m in ms], init, 0, nr, nb, seed=99,
                          reject_conc_above=a.oracle_guard if a.oracle_guard > 0 else None)
t_orc = time.time() - t0
used = d[:, :, nb:, :]
lp_o = st_o["target_log_prob"][:, :, nb:]
sane = lp_o.max(axis=2) < np.median(lp_o, axis=(1, 2))[:, None] + 500.0       # oracle chains outside the precision-loss mode
keep = sane.sum(1) >= 2
K1 = K - 1
sig = used[..., :D, None]                                                      # [B,C,S,D,1]
boff = used[..., D:D + D * K1].reshape(used.shape[:3] + (D, K1))
iraw = used[..., D + D * K1:D + 2 * D * K1].reshape(used.shape[:3] + (D, K1))
with np.errstate(over="ignore"):
    beta = 1.0 / (1.0 + np.exp(-50.0 * iraw)) * sig * boff                     # [B,C,S,D,K-1]; the reference column is last
nzm = np.abs(beta) > 1e-3
inc_chain = nzm.mean(2)                                                        # [B,C,D,K-1]
w4 = sane[:, :, None, None]
wgt = w4 / np.maximum(sane.sum(1), 1)[:, None, None, None]
pad = np.zeros((B, D, 1))
inc_o = np.concatenate([(inc_chain * wgt).sum(1), pad], axis=2).reshape(B, D * K)
nz = (np.where(nzm, beta, 0.0).sum(2) * w4).sum(1) / np.maximum((nzm.sum(2) * w4).sum(1), 1)
nz_o = np.concatenate([nz, pad], axis=2).reshape(B, D * K)
inc_d = res["inclusion_prob"]
mean_c = (inc_chain * wgt).sum(1, keepdims=True)
var_c = (((inc_chain - mean_c) ** 2) * w4).sum(1) / np.maximum(sane.sum(1)[:, None, None] - 1, 1)
se = float(np.sqrt(np.mean((var_c / np.maximum(sane.sum(1), 1)[:, None, None])[keep])))
nonref = (np.arange(D * K) % K) != K - 1
diff = (inc_d - inc_o)[keep][:, nonref]
thr_o, final_o = sch.credible_effect_calls(inc_o, nz_o, 0.05)
call_d, call_o = (res["final_effect"] != 0)[keep], (final_o != 0)[keep]
wt = w_true.reshape(B, D * K)[keep]
true_mask = wt != 0
strong_mask = wt >= 1.0
# intercepts and effects: posterior means pooled over the chains; Monte-Carlo error from the spread of the oracle's
# chain means.  `Final Parameter` of the intercepts = the 3-decimal rounded mean (result_classes.py:199-206).
alpha_chain = used[..., -K:].mean(2)                                            # [B,C,K]
alpha_o = (alpha_chain * sane[:, :, None]).sum(1) / np.maximum(sane.sum(1), 1)[:, None]
alpha_d = res["mean"][:, -K:]
se_alpha = float(np.sqrt(np.mean((alpha_chain.var(axis=1, ddof=1) / C)[keep])))
d_alpha = (alpha_d - alpha_o)[keep]
beta_chain = beta.mean(2)                                                       # [B,C,D,K-1]
beta_o = (beta_chain * w4).sum(1) / np.maximum(sane.sum(1), 1)[:, None, None]
import torch
beta_d = shard.problem.split_summary(shard.summary)["beta_mean"].mean(1).cpu().numpy().reshape(B, D, K)[:, :, :K1]
se_beta = float(np.sqrt(np.mean((beta_chain.var(axis=1, ddof=1) / C)[keep])))
d_beta = (beta_d - beta_o)[keep]
both_called = ((res["final_effect"] != 0) & (final_o != 0))[keep]
d_final = (res["final_effect"] - final_o)[keep][both_called]

# posterior mass above the fp32 guard: the device in fp64 (guard 1e15), largest concentration of every retained draw
nb64 = min(a.fp64_datasets, B)
cfg64 = sch.SweepConfig(method=0, num_results=nr, num_burnin=nb, chains=C, seed=13, dtype=np.float64)
res64, shard64 = sch.run_sweep(x[:nb64], y[:nb64], K - 1, cfg64)
d64 = shard64.out.draws[:, :, nb:, :]                                           # [b,C,S,P] fp64 on the device
sig64 = d64[..., :D].unsqueeze(-1)
boff64 = d64[..., D:D + D * K1].reshape(d64.shape[:3] + (D, K1))
iraw64 = d64[..., D + D * K1:D + 2 * D * K1].reshape(d64.shape[:3] + (D, K1))
beta64 = torch.sigmoid(50.0 * iraw64) * sig64 * boff64                          # [b,C,S,D,K-1]
xu = torch.as_tensor(np.unique(x[:nb64].reshape(-1, D), axis=0), device=d64.device)          # distinct covariate rows
eta = d64[..., -K:].unsqueeze(-2).repeat(1, 1, 1, xu.shape[0], 1)               # [b,C,S,U,K]
eta[..., :K1] += torch.einsum("ud,bcsdk->bcsuk", xu, beta64)
max_conc = torch.exp(eta.amax(dim=(-1, -2)))                                    # [b,C,S]
mass = {"datasets": int(nb64), "draws": int(max_conc.numel()), "kernel_kind": int(shard64.out.kernel_kind),
        "max_concentration": float(max_conc.max().item()),
        "share_of_draws_with_a_concentration_above_1e6": float((max_conc > 1e6).double().mean().item()),
        "share_above_1e4": float((max_conc > 1e4).double().mean().item()),
        "quantiles_50_99_9999": [float(v) for v in torch.quantile(max_conc.flatten()[::7].double(), torch.tensor([0.5, 0.99, 0.9999], dtype=torch.float64, device=d64.device)).cpu()]}

out = {
    "alpha_mean_rms_diff": float(np.sqrt(np.mean(d_alpha ** 2))), "alpha_mean_max_abs_diff": float(np.abs(d_alpha).max()),
    "alpha_mean_standard_error": se_alpha, "intercept_final_parameter_max_abs_diff": float(np.abs(np.round(alpha_d, 3) - np.round(alpha_o, 3))[keep].max()),
    "beta_mean_rms_diff": float(np.sqrt(np.mean(d_beta ** 2))), "beta_mean_max_abs_diff": float(np.abs(d_beta).max()),
    "beta_mean_standard_error": se_beta,
    "effect_final_parameter_rms_diff_where_both_call": float(np.sqrt(np.mean(d_final ** 2))) if d_final.size else None,
    "effect_final_parameter_max_abs_diff_where_both_call": float(np.abs(d_final).max()) if d_final.size else None,
    "oracle_guard": a.oracle_guard, "posterior_mass_above_fp32_guard_from_fp64_device_run": mass,
    "workload": f"{B} datasets x {C} chains, K={K}, N={y.shape[1]}, D={D}, sample_hmc({nr},{nb}), L=10; device fp32 (kernel kind {shard.out.kernel_kind}) vs C oracle fp64, independent streams",
    "device_wall_s": t_dev, "device_kernel_ms": shard.out.kernel_ms, "oracle_wall_s": t_orc,
    "oracle_chains_in_precision_loss_mode": float(1.0 - sane.mean()), "datasets_compared": int(keep.sum()),
    "entries": int(diff.size), "typical_standard_error_of_an_entry": se,
    "inclusion_prob_rms_diff": float(np.sqrt(np.mean(diff ** 2))), "inclusion_prob_max_abs_diff": float(np.abs(diff).max()),
    "inclusion_prob_mean_diff": float(diff.mean()),
    "inclusion_prob_correlation": float(np.corrcoef(inc_d[keep][:, nonref].ravel(), inc_o[keep][:, nonref].ravel())[0, 1]),
    "credible_call_disagreement_rate": float((call_d != call_o).mean()),
    "credible_calls_device": int(call_d.sum()), "credible_calls_oracle": int(call_o.sum()),
    "true_effect_called_device": float(call_d[true_mask].mean()), "true_effect_called_oracle": float(call_o[true_mask].mean()),
    "true_effect_called_device_strong": float(call_d[strong_mask].mean()) if strong_mask.any() else None,
    "true_effect_called_oracle_strong": float(call_o[strong_mask].mean()) if strong_mask.any() else None,
    "true_effect_same_call": float((call_d[true_mask] == call_o[true_mask]).mean()),
    "false_positive_rate_device": float(call_d[~true_mask & nonref[None, :]].mean()),
    "false_positive_rate_oracle": float(call_o[~true_mask & nonref[None, :]].mean()),
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
