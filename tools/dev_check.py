"""Development helper (run on the GPU box): prints K1 parity errors and quick HMC numbers."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests")); sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from conftest import haber_salm, random_states, synthetic_dataset
from oracle import np_oracle as o, c_oracle as co
from sccoda_b200.engine import DeviceProblem
from sccoda_b200 import _native as nat
import torch

for dtype in (np.float64, np.float32):
    for (N, K, D) in [(6, 8, 1), (20, 20, 1), (200, 50, 4)]:
        if N == 6:
            x, y, _ = haber_salm()
        else:
            x, y = synthetic_dataset(7, N=N, K=K, D=D)
        rs = np.random.RandomState(0)
        th = random_states(rs, D, K, 8)
        m = o.prepare(x, y, K - 1)
        lp_ref = np.array([o.logp(t, m) for t in th]); g_ref = np.array([o.grad(t, m) for t in th])
        prob = DeviceProblem.from_arrays(x, y, K - 1, chains=8, dtype=dtype)
        lp, g = prob.logp_grad(th[None])
        lp = lp.cpu().numpy()[0]; g = g.double().cpu().numpy()[0]
        print(dtype.__name__, (N, K, D), "U", prob.packed.Umax, "lp abs err", np.abs(lp - lp_ref).max(), "rel", (np.abs(lp - lp_ref) / np.abs(lp_ref)).max(),
              "grad rel err", (np.abs(g - g_ref) / np.abs(g_ref).max(1, keepdims=True)).max())

# HMC short trajectory parity fp64
x, y, _ = haber_salm()
m = o.prepare(x, y, 5)
rs = np.random.RandomState(0)
D, K = 1, 8
init = np.concatenate([np.ones(D), rs.randn(D * (K - 1)), np.zeros(D * (K - 1)), rs.randn(K)])
for method in (0, 1, 2):
    nr = 60
    dc, sc = co.sample(m.x[None], m.y[None], m.n_total[None], [m.logp_const], [m.ref], init[None, None], method, nr, 0, num_adapt=40, seed=7, chain_id0=3)
    prob = DeviceProblem.from_arrays(x, y, 5, chains=1, dtype=np.float64)
    smp = prob.make_sampler(method, nr, 0, num_adapt=40, seed=7, chain_id0=3)
    out = prob.sample(init[None, None], smp)
    dg = out.draws.cpu().numpy()[0, 0]
    err = np.abs(dg - dc[0, 0]).max(1)
    print("method", method, "W", out.warps_per_chain, "err@[0,5,10,20,40]", err[[0, 5, 10, 20, 40]], "ms", out.kernel_ms)
    print("   leap gpu", out.stat("leapfrogs_taken").cpu().numpy()[0, 0, :10], "oracle", sc["leapfrogs_taken"][0, 0, :10])

# throughput quick look: cfg5-like
B, C, N, K, D = 1024, 4, 20, 20, 1
xs, ys = zip(*[synthetic_dataset(1234 + b, N=N, K=K, D=D) for b in range(B)])
from sccoda_b200._pack import pack
pk = pack(np.stack(xs), np.stack(ys), K - 1, dtype=np.float32)
rs = np.random.RandomState(1)
P = pk.P
init = np.zeros((B, C, P), np.float32); init[..., :D] = 1; init[..., D:D + D * (K - 1)] = rs.randn(B, C, D * (K - 1)); init[..., -K:] = rs.randn(B, C, K)
prob = DeviceProblem(pk, chains=C)
for W in (1, 2, 4):
    smp = prob.make_sampler(0, 1000, 250, seed=1, warps_per_chain=W)
    out = prob.sample(init, smp)
    out = prob.sample(init, smp)
    ge = B * C * 1250 * 10
    print("cfg5 W", W, "cpc", out.chains_per_cta, "grid", out.grid, "smem", out.smem_bytes, "ms", out.kernel_ms, "grad-evals/s %.3e" % (ge / out.kernel_ms * 1e3),
          "acc", out.stat("is_accepted").mean().item(), "eps", out.stat("step_size")[..., -1].mean().item())
