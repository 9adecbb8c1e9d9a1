"""The transition of tests/golden/trapped_state.npz on the device (tools, not product): the oracle chain entered the
precision-loss mode at traced step `step_in` from `theta_before`.  Resume the device chain (same seed / chain id, same
step index, same step size) at that state and see what it proposes and whether it accepts — fp64 and fp32, guard on/off."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import c_oracle, np_oracle as o
from sccoda_b200 import _native as nat
from sccoda_b200 import scheduler as sch
from sccoda_b200._pack import pack
from sccoda_b200.engine import DeviceProblem
from sccoda_b200.util.data_generation import generate_sweep

z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "trapped_state.npz"))
b, c, K = int(z["dataset"]), int(z["chain"]), 20
x, y, _ = generate_sweep(b + 1, K=K, n_per_group=10, n_total=5000, seed0=1234)
m = o.prepare(x[b], y[b], K - 1)
# the oracle's own trace around the event (step size of the step, draws)
init = sch.initial_states(12, 4, 1, K, seed=99)[b:b + 1]
d, st = c_oracle.sample(m.x[None], m.y[None], m.n_total[None], [m.logp_const], [19], init, 0, 20000, 5000, seed=99, chain_id0=4 * b)
t_in = int(z["step_in"])
eps = st["step_size"][0, c, t_in]
print("oracle: step", t_in, "eps", eps, "lp before", st["target_log_prob"][0, c, t_in - 1], "lp at", st["target_log_prob"][0, c, t_in],
      "accepted", st["is_accepted"][0, c, t_in], "log_accept_ratio", st["log_accept_ratio"][0, c, t_in])
P = m.P
for dtype in (np.float64, np.float32):
    for guard in (0.0, -1.0):
        prob = DeviceProblem(pack(x[b], y[b], K - 1, dtype=dtype), chains=1)
        smp = prob.make_sampler(0, 20000, 5000, seed=99, chain_id0=4 * b + c, reject_conc_above=guard)
        carry = torch.zeros((1, 1, P + nat.NCARRY), dtype=torch.float64, device="cuda")
        carry[0, 0, :P] = torch.as_tensor(d[0, c, t_in - 1])
        carry[0, 0, P + 0] = float(eps)
        draws = torch.zeros((1, 1, 20000, P), dtype=prob.tdtype, device="cuda")
        stats = torch.zeros((1, 1, 20000, nat.NSTAT), dtype=prob.tdtype, device="cuda")
        smp.step_begin, smp.step_end = 5000 + t_in, 5000 + t_in + 3
        import ctypes
        nat.check(nat.lib().sccoda_sample(ctypes.byref(prob.c_problem), ctypes.byref(smp), None, draws.data_ptr(), stats.data_ptr(),
                                          carry.data_ptr(), torch.cuda.current_stream().cuda_stream), "sccoda_sample")
        torch.cuda.synchronize()
        s = stats[0, 0, t_in:t_in + 3].double().cpu().numpy()
        dd = draws[0, 0, t_in].double().cpu().numpy()
        print(f"device {np.dtype(dtype).name} guard {'default' if guard == 0 else 'off'}: accepted {s[:, 2]} lar {s[:, 1]} lp {s[:, 0]} "
              f"|draw - oracle theta_first|max {np.abs(dd - z['theta_first']).max():.3g} |draw - theta_before|max {np.abs(dd - z['theta_before']).max():.3g}")
