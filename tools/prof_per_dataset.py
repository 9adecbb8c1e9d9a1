"""Latency of one chain per dataset (tools, not product): each selected dataset of the sweep is replicated into a launch of
its own that under-fills the device (one warp per scheduler at most), so the kernel time is that dataset's chain latency."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sccoda_b200 import scheduler as sch
from sccoda_b200._pack import pack
from sccoda_b200.engine import DeviceProblem
from sccoda_b200.util.data_generation import generate_sweep

n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
x, y, _ = generate_sweep(128, K=20, n_per_group=10)
res = []
for b in range(n):
    xx, yy = np.repeat(x[b:b + 1], 32, 0), np.repeat(y[b:b + 1], 32, 0)
    pk = pack(xx, yy, 19, dtype=np.float32)
    prob = DeviceProblem(pk, chains=4)
    init = sch.initial_states(32, 4, 1, 20, seed=1).astype(np.float32)
    smp = prob.make_sampler(0, 2000, 500, step_size=0.02, seed=3)
    out = prob.sample(init, smp)
    out = prob.sample(init, smp)
    small = int((pk.y[0] < 6).sum())
    zeros = int((y[b] == 0).sum())
    res.append((out.kernel_ms, pk.NT, pk.NOT, small, zeros))
    print(f"dataset {b}: {out.kernel_ms:.2f} ms  NT={pk.NT} NOT={pk.NOT} counts<6={small} zeros={zeros}")
ms = np.array([r[0] for r in res])
print("min / median / max ms:", ms.min(), np.median(ms), ms.max())
