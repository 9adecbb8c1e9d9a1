"""sample_nuts on Haber/Salmonella through the engine: generic NUTS kernel vs the case/control NUTS kernel (tools)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import haber_salm
from sccoda_b200.engine import DeviceProblem
x, y, _ = haber_salm()
rs = np.random.RandomState(0)
for C in (1, 4, 64, 4096):
    init = np.zeros((1, C, 23)); init[..., 0] = 1; init[..., 1:8] = rs.randn(1, C, 7); init[..., -8:] = rs.randn(1, C, 8)
    prob = DeviceProblem.from_arrays(x, y, 5, chains=C, dtype=np.float32)
    for generic in (True, False):
        nr, nb = (2000, 500) if C < 1000 else (300, 200)
        smp = prob.make_sampler(2, nr, nb, seed=3, generic_only=generic)
        out = prob.sample(init, smp)
        out = prob.sample(init, smp)
        lf = out.stat("leapfrogs_taken").double()
        ge = float(lf.sum().item()) * (nr + nb) / nr
        print(f"chains={C:5d} kind={out.kernel_kind} W={out.warps_per_chain} ms={out.kernel_ms:9.1f} leapfrogs/draw={lf.mean().item():6.1f} "
              f"grad-evals/s={ge / out.kernel_ms * 1e3:.3e}")
