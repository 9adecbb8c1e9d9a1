"""NUTS on one dataset with a continuous covariate (one covariate group per row, element layout) — the generic NUTS
kernel with many warps per chain (tools)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sccoda_b200 import scheduler as sch
from sccoda_b200._pack import pack
from sccoda_b200.engine import DeviceProblem
from sccoda_b200.util.data_generation import generate_sweep
for N, K, C, seed in ((50, 10, 4, 3), (100, 20, 4, 3), (100, 20, 4, 4), (100, 20, 4, 5), (60, 40, 1, 3)):
    x, y, _ = generate_sweep(1, K=K, n_per_group=N // 2)
    x = np.random.RandomState(0).randn(1, N, 1)
    pk = pack(x, y, K - 1, dtype=np.float32)
    prob = DeviceProblem(pk, chains=C)
    init = sch.initial_states(1, C, 1, K, seed=1).astype(np.float32)
    out = prob.sample(init, prob.make_sampler(2, 300, 200, seed=seed))
    lf = out.stat("leapfrogs_taken").double()
    ge = float(lf.sum().item()) * 500 / 300
    print(f"N={N} K={K} chains={C} seed={seed}: kind={out.kernel_kind} W={out.warps_per_chain} grouped={pk.grouped} ms={out.kernel_ms:.1f} "
          f"leapfrogs/draw={lf.mean().item():.1f} grad-evals/s={ge / out.kernel_ms * 1e3:.3e}")
