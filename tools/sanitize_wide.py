"""Small launches of the new code paths of round 2 for compute-sanitizer (tools): the wide geometry (cc and mg), the
under-filled entry points, the odd-count staging with extra items, the family logp/grad kernels, the raw-host call."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import synthetic_dataset
from sccoda_b200._pack import pack
from sccoda_b200.engine import DeviceProblem, sample_raw_host
from sccoda_b200.util.data_generation import generate_sweep

rs = np.random.RandomState(0)
for B, C in ((700, 4), (400, 4), (40, 4), (2500, 1)):
    x, y, _ = generate_sweep(B, K=20, n_per_group=10, n_total=5000, seed0=5)
    y[0, :7, 0] = 0
    y[1, 10:, 1] = [0, 0.25, 2.5, 5.75, 0, 0.25, 3, 1, 0, 7]
    pk = pack(x, y, np.arange(B) % 20, dtype=np.float32)
    prob = DeviceProblem(pk, chains=C)
    init = np.zeros((B, C, pk.P), np.float32); init[..., 0] = 1; init[..., 1:20] = rs.randn(B, C, 19); init[..., -20:] = rs.randn(B, C, 20)
    out = prob.sample(init, prob.make_sampler(1, 6, 4, seed=1), chunk=7)
    print("cc", B, C, "plan", prob.plan(prob.make_sampler(1, 6, 4, seed=1)), "kind", out.kernel_kind, flush=True)
    if B <= 400:
        prob.logp_grad(init, kernel_kind=3)
        prob.logp_grad(init, kernel_kind=5)
N = 15
_, y, _ = generate_sweep(520, K=20, n_per_group=8, n_total=5000, seed0=9)
y = y[:, :N]; y[1, 2, 3] = 0
lv = np.arange(N) % 3
x = np.broadcast_to((lv[:, None] == np.arange(1, 3)[None, :]).astype(float), (520, N, 2)).copy()
pk = pack(x, y, 19, dtype=np.float32)
prob = DeviceProblem(pk, chains=4)
init = np.zeros((520, 4, pk.P), np.float32); init[..., :2] = 1; init[..., -20:] = rs.randn(520, 4, 20)
out = prob.sample(init, prob.make_sampler(0, 6, 4, seed=2))
print("mg plan", prob.plan(prob.make_sampler(0, 6, 4, seed=2)), "kind", out.kernel_kind, flush=True)
xs, ys, _ = generate_sweep(8, K=20, n_per_group=10, seed0=3)
sample_raw_host(xs, ys, np.full(8, 19, np.int32), 4, np.ascontiguousarray(init[:8, :, :59] * 0 + 0.1), prob.make_sampler(0, 6, 4, seed=3), want_draws=True, want_summary=True, summary_skip=2)
print("sanitize wide done")
