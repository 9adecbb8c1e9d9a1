"""Tiny run of every kernel (K1, K2 simple + DA, K3, K4, host entry) for compute-sanitizer (tools, not product)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import haber_salm, synthetic_dataset
from sccoda_b200 import _native as nat
from sccoda_b200._pack import pack
from sccoda_b200.engine import DeviceProblem, sample_host

x, y, _ = haber_salm()
rs = np.random.RandomState(0)
for dtype in (np.float32, np.float64):
    for (xx, yy, ref, C) in ((x, y, 5, 3), (*synthetic_dataset(3, N=12, K=9, D=3), 2, 2)):
      for grouped in (False, True):
        pk = pack(xx, yy, ref, dtype=dtype, grouped=grouped)
        prob = DeviceProblem(pk, chains=C)
        init = np.zeros((1, C, pk.P)); init[..., :pk.D] = 1; init[..., pk.D:] = 0.3 * rs.randn(1, C, pk.P - pk.D)
        init[..., pk.D + pk.D * (pk.K - 1):pk.D + 2 * pk.D * (pk.K - 1)] = 0
        prob.logp_grad(init)
        for method, W in ((0, 1), (1, 2), (2, 1), (2, 4)):
            out = prob.sample(init, prob.make_sampler(method, 12, 6, seed=1, warps_per_chain=W, max_tree_depth=5), chunk=7)
            prob.summarize(out.draws, out.stats, 3)
        sample_host(pk, C, init.astype(dtype), prob.make_sampler(0, 10, 5, seed=2), want_draws=True, want_summary=True)
# the case/control kernel (fp32, one warp per chain, grouped layout), with pseudocounts and an odd number of chains
xs, ys = zip(*[synthetic_dataset(20 + b, N=10, K=6) for b in range(3)])
ys = np.stack(ys); ys[0, 1, 2] = 0; ys[2, 3, 0] = 0; ys[2, 4, 0] = 0
pk = pack(np.stack(xs), ys, np.array([5, 0, 2]), dtype=np.float32, grouped=True)
prob = DeviceProblem(pk, chains=3)
init = np.zeros((3, 3, pk.P)); init[..., :1] = 1; init[..., 1:] = 0.3 * rs.randn(3, 3, pk.P - 1); init[..., 6:11] = 0
for method in (0, 1):
    out = prob.sample(init, prob.make_sampler(method, 12, 6, seed=1, warps_per_chain=1), chunk=7)
    assert out.kernel_kind == 3, out.kernel_kind
    prob.summarize(out.draws, out.stats, 3)
print("sanitize smoke done")
