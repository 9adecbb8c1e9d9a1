"""Where does a single-dataset `sample_*()` call spend its wall time, and which warps-per-chain setting gives the
lowest single-chain latency?  (tool, not product)"""
import cProfile
import io
import os
import pstats
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sccoda_b200 import _native as nat
from sccoda_b200 import datasets, scheduler as sch
from sccoda_b200._pack import pack
from sccoda_b200.engine import DeviceProblem
from sccoda_b200.util import cell_composition_data as dat, comp_ana as mod

df = datasets.haber()
d = dat.from_pandas(df, covariate_columns=["Mouse"])
d.obs["Condition"] = d.obs["Mouse"].str.replace(r"_[0-9]", "", regex=True)
model = mod.CompositionalAnalysis(d, formula="Condition", reference_cell_type="Goblet")
model.sample_hmc(num_results=200, num_burnin=50, verbose=False, seed=1)      # CUDA init, library load

for name, call in (("sample_hmc", lambda: model.sample_hmc(verbose=False, seed=2)),
                   ("sample_hmc_da x4", lambda: model.sample_hmc_da(verbose=False, seed=2, num_chains=4)),
                   ("sample_nuts", lambda: model.sample_nuts(verbose=False, seed=2))):
    pr = cProfile.Profile()
    t0 = time.time()
    pr.enable()
    res = call()
    pr.disable()
    wall = time.time() - t0
    t1 = time.time()
    res.summary_prepare()
    t_sum = time.time() - t1
    print(f"== {name}: wall {wall:.3f} s, kernel {model._last_run['kernel_ms']:.1f} ms, W {model._last_run['warps_per_chain']}, "
          f"summary_prepare {t_sum:.3f} s", flush=True)
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(18)
    print("\n".join(s.getvalue().splitlines()[6:30]), flush=True)

# ---- single-dataset latency against warps per chain (Haber, D=3, K=8, N=10)
x = np.asarray(model.x, float)
y = np.asarray(model.y, float)
for grouped in (0, 1):
    pk = pack(x, y, int(model.reference_cell_type), dtype=np.float32, grouped=bool(grouped))
    for C in (1, 4):
        prob = DeviceProblem(pk, chains=C)
        init = sch.initial_states(1, C, model.D, model.K, seed=3).astype(np.float32)
        for method, nm in ((nat.METHOD_HMC, "hmc"), (nat.METHOD_HMC_DA, "da"), (nat.METHOD_NUTS, "nuts")):
            row = []
            for W in (1, 2, 4, 8):
                nr, nb = (20000, 5000) if method != nat.METHOD_NUTS else (4000, 1000)
                smp = prob.make_sampler(method, nr, nb, seed=5, warps_per_chain=W)
                try:
                    o = prob.sample(init, smp)
                    row.append(f"W{W} {o.kernel_ms:8.1f} ms")
                except Exception as exc:                                   # noqa: BLE001
                    row.append(f"W{W} n/a ({str(exc)[:30]})")
            print(f"grouped {grouped} chains {C} {nm:4s}: " + " | ".join(row), flush=True)
