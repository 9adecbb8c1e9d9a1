"""Continuous covariate (one covariate group per row): multi-group kernel vs the generic kernels on both layouts (tools)."""
import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sccoda_b200 import scheduler as sch
from sccoda_b200._pack import pack
from sccoda_b200.engine import DeviceProblem
from sccoda_b200.util.data_generation import generate_sweep
for N, B in ((16, 1), (32, 1), (32, 1024)):
    K, C = 20, 4
    x, y, _ = generate_sweep(B, K=K, n_per_group=N // 2)
    rs = np.random.RandomState(0)
    x = rs.randn(B, N, 1)
    init = sch.initial_states(B, C, 1, K, seed=1).astype(np.float32)
    for name, grouped, generic in (("mg", None, False), ("generic grouped", True, True), ("generic element", False, True)):
        pk = pack(x, y, K - 1, dtype=np.float32, grouped=grouped)
        prob = DeviceProblem(pk, chains=C)
        smp = prob.make_sampler(0, 400, 100, step_size=0.02, seed=3, generic_only=generic)
        out = prob.sample(init, smp)
        print(f"N={N} B={B} continuous: {name:16s} kind={out.kernel_kind} W={out.warps_per_chain} grouped={pk.grouped} ms={out.kernel_ms:.2f} grad-evals/s={B*C*500*10/out.kernel_ms*1e3:.3e}")
