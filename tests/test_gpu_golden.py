"""GPU: the CUDA path (through the C ABI) against the committed golden vectors of tests/golden/.

Tolerances (BASELINE.json north_star): 1e-5 relative in fp64 (we hold 1e-9), 1e-3 relative in fp32."""
import os

import numpy as np
import pytest

from conftest import ROOT, haber_salm

pytestmark = pytest.mark.gpu

K1 = np.load(os.path.join(ROOT, "tests", "golden", "k1_vectors.npz"))
SMP = np.load(os.path.join(ROOT, "tests", "golden", "sampler_vectors.npz"))
CASES = sorted({k.split("/")[0] for k in K1.files})


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("dtype,rel", [(np.float64, 1e-9), (np.float32, 1e-3)])
@pytest.mark.parametrize("grouped", [False, True])
def test_logp_grad_matches_golden(name, dtype, rel, grouped):
    from sccoda_b200.engine import DeviceProblem
    x, y, ref, th = K1[name + "/x"], K1[name + "/y"], int(K1[name + "/ref"]), K1[name + "/theta"]
    lp_ref, g_ref = K1[name + "/logp"], K1[name + "/grad"]
    prob = DeviceProblem.from_arrays(x, y, ref, chains=len(th), dtype=dtype, grouped=grouped)
    lp, g = prob.logp_grad(th[None])
    lp, g = lp.cpu().numpy()[0], g.double().cpu().numpy()[0]
    assert np.all(np.abs(lp - lp_ref) <= rel * np.maximum(1.0, np.abs(lp_ref))), np.abs(lp - lp_ref).max()
    scale = np.abs(g_ref).max(axis=1, keepdims=True)
    assert (np.abs(g - g_ref) / scale).max() <= rel


@pytest.mark.parametrize("method,steps,tol", [(0, 30, 1e-9), (1, 8, 1e-7), (2, 6, 1e-6)])
@pytest.mark.parametrize("grouped", [False, True])
def test_fp64_samplers_walk_the_golden_trajectories(method, steps, tol, grouped):
    from sccoda_b200.engine import DeviceProblem
    x, y, _ = haber_salm()
    prob = DeviceProblem.from_arrays(x, y, 5, chains=1, dtype=np.float64, grouped=grouped)
    smp = prob.make_sampler(method, steps, 0, num_adapt=steps, seed=7, chain_id0=3)
    out = prob.sample(SMP["init"][None, None], smp)
    assert np.abs(out.draws.cpu().numpy()[0, 0] - SMP[f"m{method}/draws"]).max() < tol
    for k in ("is_accepted", "leapfrogs_taken", "diverging", "reach_max_depth"):
        assert np.array_equal(out.stat(k).cpu().numpy()[0, 0], SMP[f"m{method}/{k}"]), k
