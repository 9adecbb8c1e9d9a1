"""Per-chain comparison of device vs oracle on the sweep datasets (debugging aid, tools)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import c_oracle, np_oracle as o
from sccoda_b200 import scheduler as sch
from sccoda_b200._pack import pack
from sccoda_b200.engine import DeviceProblem
from sccoda_b200.util.data_generation import generate_sweep

B, C, K, nr, nb = 12, 16, 20, 6000, 2000
x, y, w_true = generate_sweep(B, K=K, n_per_group=10)
init = sch.initial_states(B, C, 1, K, seed=5)
ms = [o.prepare(x[b], y[b], K - 1) for b in range(B)]
d_o, s_o = c_oracle.sample(np.stack([m.x for m in ms]), np.stack([m.y for m in ms]), np.stack([m.n_total for m in ms]),
                           [m.logp_const for m in ms], [m.ref for m in ms], init, 0, nr, nb, seed=99)
outs = {}
for name, kw in (("cc", {}), ("generic", dict(generic_only=True)), ("f64", dict())):
    dt = np.float64 if name == "f64" else np.float32
    prob = DeviceProblem(pack(x, y, K - 1, dtype=dt), chains=C)
    out = prob.sample(init, prob.make_sampler(0, nr, nb, seed=11, **kw))
    outs[name] = (out.draws.double().cpu().numpy(), out.stat("is_accepted").double().cpu().numpy(), out.stat("step_size").double().cpu().numpy(), out.kernel_kind)

def stats(d):
    used = d[:, :, nb:, :]
    sig, boff, iraw = used[..., 0:1], used[..., 1:K], used[..., K:2 * K - 1]
    beta = 1.0 / (1.0 + np.exp(-50.0 * iraw)) * sig * boff
    inc = (np.abs(beta) > 1e-3)
    return inc.mean(2), np.where(inc, beta, 0).sum(2) / np.maximum(inc.sum(2), 1)

inc_o, nz_o = stats(d_o)
print("oracle acc", s_o["is_accepted"].mean(), "eps", s_o["step_size"][:, :, -1].mean())
for name, (d, acc, eps, kind) in outs.items():
    inc, nz = stats(d)
    print(name, "kind", kind, "acc", acc.mean(), "eps", eps[:, :, -1].mean())
    for b in range(B):
        k = int(np.argmax(w_true[b]))
        print(f"  b={b} k={k} w={w_true[b,k]:.1f}  inc dev {inc[b,:,k].mean():.3f}+-{inc[b,:,k].std()/4:.3f} orc {inc_o[b,:,k].mean():.3f}+-{inc_o[b,:,k].std()/4:.3f}"
              f"   nz dev {nz[b,:,k].mean():.3f}+-{nz[b,:,k].std()/4:.3f} orc {nz_o[b,:,k].mean():.3f}+-{nz_o[b,:,k].std()/4:.3f}"
              f"   alpha_k dev {d[b,:,nb:,-K+k].mean():.3f} orc {d_o[b,:,nb:,-K+k].mean():.3f}")
