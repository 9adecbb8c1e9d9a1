"""Device time of the single-dataset cases users actually run (tools): Haber Salmonella (D = 1) and all four conditions
(D = 3), 1 and 4 chains, sample_hmc / sample_hmc_da / sample_nuts at their default lengths."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sccoda_b200 import _native as nat
from sccoda_b200 import datasets, scheduler as sch
from sccoda_b200._pack import pack
from sccoda_b200.engine import DeviceProblem

df = datasets.haber()
y_all = df.iloc[:, 1:].to_numpy(float)
cond = df["Mouse"].str.replace(r"_[0-9]", "", regex=True)
lv = sorted(cond.unique())
x3 = np.stack([(cond == l).to_numpy(float) for l in lv[1:]], axis=1)
rows = [0, 1, 2, 3, 8, 9]
cases = {"salm D=1": (np.array([0, 0, 0, 0, 1, 1.])[:, None], y_all[rows]), "all conditions D=3": (x3, y_all)}
for name, (x, y) in cases.items():
    pk = pack(x, y, 3, dtype=np.float32)
    for C in (1, 4):
        prob = DeviceProblem(pk, chains=C)
        init = sch.initial_states(1, C, x.shape[1], 8, seed=3).astype(np.float32)
        row = []
        for method, nm, nr, nb in ((nat.METHOD_HMC, "hmc", 20000, 5000), (nat.METHOD_HMC_DA, "da", 20000, 5000), (nat.METHOD_NUTS, "nuts", 10000, 5000)):
            smp = prob.make_sampler(method, nr, nb, seed=5)
            o = prob.sample(init, smp)
            o = prob.sample(init, smp)
            row.append(f"{nm} kind {o.kernel_kind} {o.kernel_ms:8.1f} ms")
        print(f"{name:20s} chains {C}: " + " | ".join(row), flush=True)
