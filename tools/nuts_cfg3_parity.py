"""BASELINE config 3 (N = 200, K = 50, D = 4, sample_nuts, depth 10): device fp32 against the C oracle fp64 on chains long
enough to compare statistics (tools, not product).  Writes one JSON object; the oracle needs minutes.

    python tools/nuts_cfg3_parity.py [chains] [num_results] [num_burnin]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_sampler import _config3_data, _init, _oracle_run       # noqa: E402
from sccoda_b200.engine import DeviceProblem                        # noqa: E402

C = int(sys.argv[1]) if len(sys.argv) > 1 else 16
nr = int(sys.argv[2]) if len(sys.argv) > 2 else 150
nb = int(sys.argv[3]) if len(sys.argv) > 3 else 250
x, y = _config3_data()
K, D = 50, 4
init = _init(np.random.RandomState(4), D, K, C)
init[:, -K:] = np.log(y.mean(0) / y.mean(0).sum() * 50.0) + 0.1 * init[:, -K:]
t0 = time.time()
d_ref, s_ref = _oracle_run(x, y, K - 1, init, 2, nr, nb, seed=41, chain_id0=0)
t_orc = time.time() - t0
prob = DeviceProblem.from_arrays(x, y, K - 1, chains=C, dtype=np.float32)
out = prob.sample(init[None], prob.make_sampler(2, nr, nb, seed=42))
d = out.draws.double().cpu().numpy()
st = lambda n: out.stat(n).double().cpu().numpy()[0]
am, ar = d[0, :, :, -K:].mean(1), d_ref[0, :, :, -K:].mean(1)
se = np.hypot(am.std(0, ddof=1), ar.std(0, ddof=1)) / np.sqrt(C)
K1 = K - 1
beta = lambda dd: 1.0 / (1.0 + np.exp(-50.0 * dd[..., D + D * K1:D + 2 * D * K1])) * np.repeat(dd[..., :D], K1, axis=-1) * dd[..., D:D + D * K1]
bm, br = beta(d[0]).mean(1), beta(d_ref[0]).mean(1)
seb = np.hypot(bm.std(0, ddof=1), br.std(0, ddof=1)) / np.sqrt(C)
res = {
    "workload": f"N=200 K=50 D=4, {C} chains, sample_nuts({nr},{nb}) depth 10; device fp32 (kind {out.kernel_kind}, W={out.warps_per_chain}) vs C oracle fp64, independent streams",
    "device_kernel_ms": out.kernel_ms, "oracle_wall_s": t_orc,
    "leapfrogs_per_draw": [float(st("leapfrogs_taken").mean()), float(s_ref["leapfrogs_taken"][0].mean())],
    "mean_accept_statistic": [float(np.exp(np.minimum(st("log_accept_ratio"), 0)).mean()), float(np.exp(np.minimum(s_ref["log_accept_ratio"][0], 0)).mean())],
    "final_step_size": [float(st("step_size")[:, -1].mean()), float(s_ref["step_size"][0, :, -1].mean())],
    "target_log_prob_mean": [float(st("target_log_prob").mean()), float(s_ref["target_log_prob"][0].mean())],
    "target_log_prob_sd": [float(st("target_log_prob").std()), float(s_ref["target_log_prob"][0].std())],
    "reach_max_depth_share": [float(st("reach_max_depth").mean()), float(s_ref["reach_max_depth"][0].mean())],
    "diverging_share": [float(st("diverging").mean()), float(s_ref["diverging"][0].mean())],
    "alpha_mean_max_abs_diff": float(np.abs(am.mean(0) - ar.mean(0)).max()), "alpha_mean_max_z": float((np.abs(am.mean(0) - ar.mean(0)) / se).max()),
    "alpha_mean_rms_z": float(np.sqrt(np.mean(((am.mean(0) - ar.mean(0)) / se) ** 2))),
    "beta_mean_max_abs_diff": float(np.abs(bm.mean(0) - br.mean(0)).max()),
    "beta_mean_rms_z": float(np.sqrt(np.mean(((bm.mean(0) - br.mean(0)) / np.maximum(seb, 1e-6)) ** 2))),
    "order": "[device, oracle]",
}
print(json.dumps(res, indent=1))
