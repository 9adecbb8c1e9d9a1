import sys, os, json, numpy as np
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
from test_gpu_sampler import _config3_data
import torch
from sccoda_b200 import scheduler as sch
from sccoda_b200._pack import pack
from sccoda_b200.engine import DeviceProblem
x, y = _config3_data()
K, D, C = 50, 4, 64
pk = pack(x, y, K - 1, dtype=np.float32)
init = sch.initial_states(1, C, D, K, seed=3).astype(np.float32)
prob = DeviceProblem(pk, chains=C)
nr, nb = 10000, 5000
out = prob.sample(init, prob.make_sampler(2, nr, nb, max_tree_depth=10, seed=3))
lf = out.stat("leapfrogs_taken").double()
ge = float(lf.sum().item()) * (nr + nb) / nr
d = out.draws[0, :, nb:, -K:].double()           # alpha after the second burn-in
from sccoda_b200 import diagnostics as dg
rh = dg.rhat(d.permute(2, 0, 1).contiguous()); es = dg.ess_bulk(d.permute(2, 0, 1).contiguous())
print(json.dumps(dict(config="BASELINE config 3 as written: N=200 K=50 D=4, 64 chains, sample_nuts(10000, 5000), depth 10", kernel_ms=out.kernel_ms,
                      warps_per_chain=out.warps_per_chain, grid=out.grid, mean_leapfrogs=float(lf.mean().item()), grad_evals_per_s=ge / out.kernel_ms * 1e3,
                      acc=float(torch.exp(torch.clamp(out.stat("log_accept_ratio"), max=0)).mean().item()), reach_max_depth=float(out.stat("reach_max_depth").mean().item()),
                      diverging=float(out.stat("diverging").mean().item()), rhat_alpha_max=float(rh.max().item()), ess_alpha_min=float(es.min().item()),
                      finite=bool(torch.isfinite(out.draws).all().item())), indent=1))
