"""K2/K3 parity: the persistent HMC / NUTS kernels (through the C ABI) vs the CPU oracle.

* fp64 build, same Philox streams  -> the device chain must walk the oracle's trajectory (until chaotic
  amplification of summation-order rounding, so the windows are short for the aggressive adapters).
* fp32 build -> statistical agreement over many chains (accept rate, step size, posterior means, inclusion).
"""
import numpy as np
import pytest

from conftest import haber_salm, synthetic_dataset

pytestmark = pytest.mark.gpu


def _init(rs, D, K, n):
    out = np.zeros((n, D + 2 * D * (K - 1) + K))
    out[:, :D] = 1.0                                           # scCODA_model.py:763-768
    out[:, D:D + D * (K - 1)] = rs.randn(n, D * (K - 1))
    out[:, -K:] = rs.randn(n, K)
    return out


def _oracle_run(x, y, ref, init, method, nr, nb, **kw):
    from oracle import c_oracle, np_oracle as o
    m = o.prepare(x, y, ref)
    return c_oracle.sample(m.x[None], m.y[None], m.n_total[None], [m.logp_const], [m.ref], init[None], method, nr, nb,
                           **kw)


@pytest.mark.parametrize("method,steps,tol", [(0, 40, 1e-9), (1, 8, 1e-7), (2, 6, 1e-6)])
@pytest.mark.parametrize("W", [1, 2, 4, 8, 16])      # W >= 8 + grouped + NUTS: gradient and value side by side
@pytest.mark.parametrize("grouped", [False, True])
def test_fp64_trajectory_follows_oracle(method, steps, tol, W, grouped):
    from sccoda_b200.engine import DeviceProblem
    x, y, _ = haber_salm()
    init = _init(np.random.RandomState(0), 1, 8, 3)
    d_ref, s_ref = _oracle_run(x, y, 5, init, method, steps, 0, num_adapt=steps, seed=7, chain_id0=3)
    prob = DeviceProblem.from_arrays(x, y, 5, chains=3, dtype=np.float64, grouped=grouped)
    smp = prob.make_sampler(method, steps, 0, num_adapt=steps, seed=7, chain_id0=3, warps_per_chain=W)
    out = prob.sample(init[None], smp)
    d = out.draws.cpu().numpy()
    assert np.abs(d - d_ref).max() < tol
    for name in ("leapfrogs_taken", "is_accepted", "diverging", "reach_max_depth"):
        assert np.array_equal(out.stat(name).cpu().numpy(), s_ref[name]), name
    assert np.allclose(out.stat("step_size").cpu().numpy(), s_ref["step_size"], rtol=1e-9)
    assert np.allclose(out.stat("target_log_prob").cpu().numpy(), s_ref["target_log_prob"], rtol=1e-9, atol=1e-7)
    fin = np.isfinite(s_ref["log_accept_ratio"])
    assert np.allclose(out.stat("log_accept_ratio").cpu().numpy()[fin], s_ref["log_accept_ratio"][fin], atol=1e-6)


@pytest.mark.parametrize("grouped", [False, True])
def test_fp64_trajectory_synthetic_multi_covariate(grouped):
    """D=3 covariates (U=8 groups after packing), K=9: exercises the grouped layout against the plain oracle."""
    from sccoda_b200.engine import DeviceProblem
    x, y = synthetic_dataset(5, N=12, K=9, D=3)
    y[0, 0] = 0
    init = _init(np.random.RandomState(1), 3, 9, 2)
    d_ref, s_ref = _oracle_run(x, y, 2, init, 0, 30, 10, num_adapt=30, seed=5, chain_id0=0)
    prob = DeviceProblem.from_arrays(x, y, 2, chains=2, dtype=np.float64, grouped=grouped)
    out = prob.sample(init[None], prob.make_sampler(0, 30, 10, num_adapt=30, seed=5))
    assert np.abs(out.draws.cpu().numpy() - d_ref).max() < 1e-8


def test_chunked_resume_is_bit_identical():
    from sccoda_b200.engine import DeviceProblem
    x, y, _ = haber_salm()
    init = _init(np.random.RandomState(2), 1, 8, 2)
    for method, nr, nb in ((0, 60, 40), (1, 60, 40), (2, 12, 8)):
        prob = DeviceProblem.from_arrays(x, y, 5, chains=2, dtype=np.float64)
        one = prob.sample(init[None], prob.make_sampler(method, nr, nb, seed=3))
        many = prob.sample(init[None], prob.make_sampler(method, nr, nb, seed=3), chunk=7)
        assert many.launches > 1
        assert np.array_equal(one.draws.cpu().numpy(), many.draws.cpu().numpy()), method
        assert np.array_equal(one.stats.cpu().numpy(), many.stats.cpu().numpy(), equal_nan=True), method


def test_chain_streams_are_independent_of_sharding():
    """chain_id0 offsets: running chains [0,4) at once equals running [0,2) and [2,4) separately (multi-GPU sharding)."""
    from sccoda_b200.engine import DeviceProblem
    x, y, _ = haber_salm()
    init = _init(np.random.RandomState(4), 1, 8, 4)
    full = DeviceProblem.from_arrays(x, y, 5, chains=4, dtype=np.float64)
    d_full = full.sample(init[None], full.make_sampler(0, 30, 10, seed=9)).draws.cpu().numpy()
    half = DeviceProblem.from_arrays(x, y, 5, chains=2, dtype=np.float64)
    d_a = half.sample(init[None, :2], half.make_sampler(0, 30, 10, seed=9, chain_id0=0)).draws.cpu().numpy()
    d_b = half.sample(init[None, 2:], half.make_sampler(0, 30, 10, seed=9, chain_id0=2)).draws.cpu().numpy()
    assert np.array_equal(d_full[:, :2], d_a) and np.array_equal(d_full[:, 2:], d_b)


@pytest.mark.parametrize("method", [0, 1])
@pytest.mark.parametrize("variant", ["generic", "grouped", "case_control"])
def test_fp32_statistical_parity_haber(method, variant):
    """64 fp32 device chains vs 64 fp64 oracle chains on Haber/Salmonella, sample_hmc(_da) semantics, for the
    element path, the pair/item path and the case/control kernel (one warp per chain, state in registers).

    Tolerances are Monte-Carlo tolerances: means over 64 chains; per-chain sd taken from the oracle sample."""
    from sccoda_b200.engine import DeviceProblem
    x, y, _ = haber_salm()
    C, nr, nb = 64, 3000, 1500
    init = _init(np.random.RandomState(7), 1, 8, C)
    d_ref, s_ref = _oracle_run(x, y, 5, init, method, nr, nb, seed=21, chain_id0=0)
    prob = DeviceProblem.from_arrays(x, y, 5, chains=C, dtype=np.float32, grouped=variant != "generic")
    smp = prob.make_sampler(method, nr, nb, seed=22, warps_per_chain=1 if variant == "case_control" else 0,
                            generic_only=variant != "case_control")
    out = prob.sample(init[None], smp)
    assert out.kernel_kind == (3 if variant == "case_control" else 1)
    d = out.draws.double().cpu().numpy()
    assert np.isfinite(d).all()

    def chain_stat(a):          # per-chain acceptance etc. -> mean and standard error over chains
        return a.mean(), a.std(ddof=1) / np.sqrt(len(a))

    # acceptance rate and adapted step size
    acc, acc_se = chain_stat(out.stat("is_accepted").double().cpu().numpy()[0].mean(1))
    acc_r, acc_r_se = chain_stat(s_ref["is_accepted"][0].mean(1))
    assert abs(acc - acc_r) < 5 * np.hypot(acc_se, acc_r_se) + 0.01, (acc, acc_r)
    eps, eps_se = chain_stat(out.stat("step_size").double().cpu().numpy()[0][:, -1])
    eps_r, eps_r_se = chain_stat(s_ref["step_size"][0][:, -1])
    assert abs(eps - eps_r) < 5 * np.hypot(eps_se, eps_r_se) + 0.002, (eps, eps_r)
    # posterior means of the intercepts (alpha): chain means are ~independent replicates
    K = 8
    am, ar = d[0, :, nb // 2:, -K:].mean(1), d_ref[0, :, nb // 2:, -K:].mean(1)
    se = np.hypot(am.std(0, ddof=1), ar.std(0, ddof=1)) / np.sqrt(C)
    assert np.all(np.abs(am.mean(0) - ar.mean(0)) < 5 * se + 0.02), (am.mean(0), ar.mean(0))
    # inclusion probability of the strong Enterocyte effect (k=1) and of a null effect (k=0)
    def inclusion(dr):
        sig, boff, iraw = dr[..., 0:1], dr[..., 1:8], dr[..., 8:15]
        beta = 1 / (1 + np.exp(-50 * iraw)) * sig * boff
        return (np.abs(beta) > 1e-3).mean(-2)               # [chains, K-1]
    inc, inc_r = inclusion(d[0, :, nb // 2:]), inclusion(d_ref[0, :, nb // 2:])
    se = np.hypot(inc.std(0, ddof=1), inc_r.std(0, ddof=1)) / np.sqrt(C)
    assert np.all(np.abs(inc.mean(0) - inc_r.mean(0)) < 5 * se + 0.03), (inc.mean(0), inc_r.mean(0))
    assert inc.mean(0)[1] > 0.9                              # Enterocyte is credible in every tutorial run


@pytest.mark.parametrize("method", [0, 1])
def test_case_control_kernel_follows_generic_kernel(method):
    """Same Philox streams, same arithmetic per parameter: the register-resident case/control kernel and the generic
    kernel agree step by step in fp32 until rounding differences of the summation orders are amplified."""
    from sccoda_b200._pack import pack
    from sccoda_b200.engine import DeviceProblem
    B, C, K = 3, 4, 20
    xs, ys = zip(*[synthetic_dataset(40 + b, N=20, K=K) for b in range(B)])
    ys = np.stack(ys)
    ys[1, 3, 2] = 0                                              # one dataset with a pseudocount (odd group)
    init = _init(np.random.RandomState(11), 1, K, B * C).reshape(B, C, -1)
    prob = DeviceProblem(pack(np.stack(xs), ys, np.array([19, 0, 7]), dtype=np.float32), chains=C)
    assert prob.packed.grouped == 1 and prob.packed.NF <= 4
    outs = []
    for generic_only in (True, False):
        smp = prob.make_sampler(method, 16, 0, num_adapt=16, seed=3, warps_per_chain=1, generic_only=generic_only)
        outs.append(prob.sample(init, smp))
    assert outs[0].kernel_kind == 1 and outs[1].kernel_kind == 3
    d0, d1 = outs[0].draws.cpu().numpy(), outs[1].draws.cpu().numpy()
    assert np.isfinite(d1).all()
    # dual averaging jumps to 10x the initial step size at once: its trajectories are chaotic after a few steps
    n = 8 if method == 0 else 2
    assert np.abs(d0[:, :, :n] - d1[:, :, :n]).max() < 2e-3
    assert np.array_equal(outs[0].stat("is_accepted").cpu().numpy()[:, :, :n],
                          outs[1].stat("is_accepted").cpu().numpy()[:, :, :n])
    assert np.allclose(outs[0].stat("target_log_prob").cpu().numpy()[:, :, :n],
                       outs[1].stat("target_log_prob").cpu().numpy()[:, :, :n], atol=5e-2)
    assert np.allclose(outs[0].stat("step_size").cpu().numpy()[:, :, :n],
                       outs[1].stat("step_size").cpu().numpy()[:, :, :n], rtol=1e-3)
    # the traced log-density of the case/control kernel is the oracle's log-density at the traced state
    from oracle import np_oracle as o
    lp1 = outs[1].stat("target_log_prob").cpu().numpy()
    for b in range(B):
        m = o.prepare(xs[b], ys[b], int([19, 0, 7][b]))
        for c in range(C):
            ref = np.array([o.logp(d1[b, c, i].astype(np.float64), m) for i in range(16)])
            assert np.abs(lp1[b, c] - ref).max() < 2e-3 * max(1.0, np.abs(ref).max() * 1e-3) + 4e-3, (b, c)
    # chunked resume of the case/control kernel is bit-identical to the single launch
    smp = prob.make_sampler(method, 16, 0, num_adapt=16, seed=3, warps_per_chain=1)
    many = prob.sample(init, smp, chunk=5)
    assert many.launches > 1 and np.array_equal(many.draws.cpu().numpy(), d1)


def _cc_vs_oracle(xs, ys, refs, C=2, steps=12, seed=4):
    """Run the case/control kernel and the generic kernel on a batch; return both outputs and the packed problem."""
    from oracle import np_oracle as o
    from sccoda_b200._pack import pack
    from sccoda_b200.engine import DeviceProblem
    B, K = len(xs), ys[0].shape[1]
    pk = pack(np.stack(xs), np.stack(ys), np.asarray(refs), dtype=np.float32, grouped=True)
    prob = DeviceProblem(pk, chains=C)
    init = _init(np.random.RandomState(seed), 1, K, B * C).reshape(B, C, -1) * 0.5
    init[..., 0] = 1.0
    outs = [prob.sample(init, prob.make_sampler(0, steps, 0, num_adapt=steps, seed=seed, warps_per_chain=1,
                                                generic_only=g)) for g in (True, False)]
    assert outs[0].kernel_kind == 1 and outs[1].kernel_kind == 3
    d = outs[1].draws.double().cpu().numpy()
    lp = outs[1].stat("target_log_prob").double().cpu().numpy()
    assert np.isfinite(d).all() and np.isfinite(lp).all()
    for b in range(B):
        m = o.prepare(xs[b], ys[b], int(refs[b]))
        for c in range(C):
            ref = np.array([o.logp(d[b, c, i], m) for i in range(steps)])
            assert np.abs(lp[b, c] - ref).max() < 5e-6 * np.abs(ref).max() + 4e-3, (b, c, np.abs(lp[b, c] - ref).max())
    n = min(steps, 5)
    assert np.abs(outs[0].draws.cpu().numpy()[:, :, :n] - outs[1].draws.cpu().numpy()[:, :, :n]).max() < 2e-3
    return outs, pk


@pytest.mark.parametrize("K", [2, 3, 31])
def test_case_control_kernel_cell_type_range(K):
    """Smallest and largest number of cell types the lane-per-cell-type kernel takes (K = 32 goes to the generic one)."""
    xs, ys = zip(*[synthetic_dataset(60 + b, N=10, K=K) for b in range(2)])
    _cc_vs_oracle(xs, ys, [K - 1, 0])


def test_case_control_kernel_k32_takes_the_two_columns_per_lane_kernel():
    """K = 32 no longer fits one cell type per lane plus the row totals: the multi-group kernel with two columns per lane
    takes it (K <= 62), beyond that the generic kernel."""
    from sccoda_b200.engine import DeviceProblem
    x, y = synthetic_dataset(3, N=10, K=32)
    prob = DeviceProblem.from_arrays(x, y, 31, chains=2, dtype=np.float32, grouped=True)
    out = prob.sample(_init(np.random.RandomState(0), 1, 32, 2)[None] * 0.5,
                      prob.make_sampler(0, 6, 0, num_adapt=6, seed=1, warps_per_chain=1))
    assert out.kernel_kind == 5 and np.isfinite(out.draws.cpu().numpy()).all()
    x, y = synthetic_dataset(3, N=10, K=63)
    prob = DeviceProblem.from_arrays(x, y, 31, chains=2, dtype=np.float32, grouped=True)
    out = prob.sample(_init(np.random.RandomState(0), 1, 63, 2)[None] * 0.5,
                      prob.make_sampler(0, 6, 0, num_adapt=6, seed=1, warps_per_chain=1))
    assert out.kernel_kind == 1 and np.isfinite(out.draws.cpu().numpy()).all()


@pytest.mark.parametrize("method", [0, 2])
@pytest.mark.parametrize("K,kind", [(32, "case_control"), (45, "case_control"), (62, "case_control"), (40, "conditions4"),
                                    (33, "factorial2x2"), (50, "levels3")])
def test_two_columns_per_lane_kernels_follow_generic_kernel(K, kind, method):
    """32 <= K <= 62: a lane owns the cell types l and 32 + l (hmc_mg.cuh, NS = 2).  Same checks as for K <= 31: the
    oracle's log-density at the traced states, the generic kernel's first transitions, bit-identical chunked resume."""
    from oracle import np_oracle as o
    from sccoda_b200._pack import pack
    from sccoda_b200.engine import DeviceProblem
    B, C, steps = 2, 2, 8
    if kind == "case_control":
        xs, ys = zip(*[synthetic_dataset(140 + b, N=14, K=K) for b in range(B)])
        ys = list(ys)
        ys[1][2, K - 3] = 0.0                                   # a pseudocount dataset
    else:
        xs, ys = zip(*[_multi_group_dataset(150 + b, kind, N=24, K=K, zero_frac=0.1 * (b == 1)) for b in range(B)])
    D = xs[0].shape[1]
    refs = [K - 1, 5]
    pk = pack(np.stack(xs), np.stack(ys), np.asarray(refs), dtype=np.float32)
    assert pk.grouped == 1
    prob = DeviceProblem(pk, chains=C)
    init = _init(np.random.RandomState(7), D, K, B * C).reshape(B, C, -1) * 0.4
    init[..., :D] = 1.0
    kw = dict(num_adapt=steps, seed=4, warps_per_chain=1, max_tree_depth=5)
    outs = [prob.sample(init, prob.make_sampler(method, steps, 0, generic_only=g, **kw)) for g in (True, False)]
    assert outs[0].kernel_kind == (1 if method == 0 else 2) and outs[1].kernel_kind == (5 if method == 0 else 6)
    d = outs[1].draws.double().cpu().numpy()
    lp = outs[1].stat("target_log_prob").double().cpu().numpy()
    assert np.isfinite(d).all() and np.isfinite(lp).all()
    for b in range(B):
        m = o.prepare(xs[b], ys[b], int(refs[b]))
        for c in range(C):
            want = np.array([o.logp(d[b, c, i], m) for i in range(steps)])
            assert np.abs(lp[b, c] - want).max() < 5e-6 * np.abs(want).max() + 4e-3, (b, c, np.abs(lp[b, c] - want).max())
    n = 3
    assert np.abs(outs[0].draws.cpu().numpy()[:, :, :n] - outs[1].draws.cpu().numpy()[:, :, :n]).max() < 5e-3
    for name in ("is_accepted", "leapfrogs_taken"):
        assert np.array_equal(outs[0].stat(name).cpu().numpy()[:, :, :n], outs[1].stat(name).cpu().numpy()[:, :, :n]), name
    many = prob.sample(init, prob.make_sampler(method, steps, 0, **kw), chunk=3)
    assert many.launches > 1 and np.array_equal(many.draws.cpu().numpy(), outs[1].draws.cpu().numpy())


def test_case_control_kernel_ragged_batches():
    """Datasets of one batch that differ in what the kernel's loops see: many rows per group (several items per pair),
    a dataset whose rows all share one covariate value (one group), a dataset of small counts only (no items at all for
    most pairs), pseudocounts and non-integer counts (odd counts), unbalanced groups."""
    rs = np.random.RandomState(3)
    N, K = 33, 7
    xs, ys = [], []
    for b in range(5):
        x, y = synthetic_dataset(70 + b, N=N, K=K)
        if b == 0:
            x[:] = 0.0
            x[-4:] = 1.0                                   # 29 + 4 rows: six items for the big group
        if b == 1:
            x[:] = 1.0                                     # a single covariate group
        if b == 2:
            y = rs.randint(0, 6, size=(N, K)).astype(float)            # counts 0..5 only (0 -> pseudocount)
            y[:, 0] += 1
        if b == 3:
            y[rs.rand(N, K) < 0.2] = 0.0
            y[5, 3], y[6, 3], y[7, 4] = 2.5, 0.25, 5.75
        xs.append(x)
        ys.append(y)
    outs, pk = _cc_vs_oracle(xs, ys, [6, 0, 3, 2, 5], C=3, steps=10)
    assert pk.NT >= 6 and pk.NOT >= 2 and int(pk.U.min()) == 1 and int(pk.U.max()) == 2


def _design(kind, N, rs):
    """Design matrices with more than two covariate groups."""
    if kind == "levels3":                      # one covariate with three levels: D = 1, 3 groups
        return rs.randint(0, 3, size=(N, 1)).astype(float)
    if kind == "factorial2x2":                 # D = 2, 4 groups
        return rs.randint(0, 2, size=(N, 2)).astype(float)
    if kind == "conditions4":                  # one-hot coded four conditions (the Haber design): D = 3, 4 groups
        lv = np.arange(N) % 4
        return (lv[:, None] == np.arange(1, 4)[None, :]).astype(float)
    if kind == "factorial2x2x2":               # D = 3, 8 groups
        x = ((np.arange(N)[:, None] >> np.arange(3)[None, :]) & 1).astype(float)
        return x[rs.permutation(N)]
    if kind == "conditions5":                  # D = 4, 5 groups
        lv = np.arange(N) % 5
        return (lv[:, None] == np.arange(1, 5)[None, :]).astype(float)
    if kind == "factorial2x2x2x2":             # D = 4, 16 groups
        x = ((np.arange(N)[:, None] >> np.arange(4)[None, :]) & 1).astype(float)
        return x[rs.permutation(N)]
    if kind == "groups3_d2":                   # D = 2, 3 groups (odd group count: one empty slot in the last pair)
        lv = np.arange(N) % 3
        return np.stack([lv == 1, lv == 2], axis=1).astype(float)
    raise ValueError(kind)


def _multi_group_dataset(seed, kind, N, K, zero_frac=0.0):
    rs = np.random.RandomState(seed)
    x = _design(kind, N, rs)
    D = x.shape[1]
    b = rs.uniform(-2, 2, size=K)
    w = np.zeros((D, K))
    for d in range(D):
        w[d, rs.randint(K - 1)] = rs.choice([-1.0, 0.7, 1.2])
    logits = x @ w + b + np.sqrt(0.05) * rs.randn(N, K)
    p = np.exp(logits - logits.max(1, keepdims=True))
    p /= p.sum(1, keepdims=True)
    y = np.stack([rs.multinomial(3000, p[n]) for n in range(N)]).astype(float)
    if zero_frac:
        y[rs.rand(N, K) < zero_frac] = 0.0
    return x, y


def _mg_vs_generic(xs, ys, refs, method=0, C=2, steps=12, seed=4):
    from oracle import np_oracle as o
    from sccoda_b200._pack import pack
    from sccoda_b200.engine import DeviceProblem
    B, K, D = len(xs), ys[0].shape[1], xs[0].shape[1]
    pk = pack(np.stack(xs), np.stack(ys), np.asarray(refs), dtype=np.float32)
    assert pk.grouped == 1                     # few covariate groups: the packer chooses the pair / item layout
    prob = DeviceProblem(pk, chains=C)
    init = _init(np.random.RandomState(seed), D, K, B * C).reshape(B, C, -1) * 0.5
    init[..., :D] = 1.0
    outs = [prob.sample(init, prob.make_sampler(method, steps, 0, num_adapt=steps, seed=seed, warps_per_chain=1,
                                                generic_only=g)) for g in (True, False)]
    assert outs[0].kernel_kind == 1 and outs[1].kernel_kind == 5
    d = outs[1].draws.double().cpu().numpy()
    lp = outs[1].stat("target_log_prob").double().cpu().numpy()
    assert np.isfinite(d).all() and np.isfinite(lp).all()
    # the traced log-density is the oracle's log-density at the traced state
    for b in range(B):
        m = o.prepare(xs[b], ys[b], int(refs[b]))
        for c in range(C):
            ref = np.array([o.logp(d[b, c, i], m) for i in range(steps)])
            assert np.abs(lp[b, c] - ref).max() < 5e-6 * np.abs(ref).max() + 4e-3, (b, c, np.abs(lp[b, c] - ref).max())
    # same Philox streams, same arithmetic per parameter: the first transitions agree with the generic kernel
    n = min(steps, 5) if method == 0 else 2
    assert np.abs(outs[0].draws.cpu().numpy()[:, :, :n] - outs[1].draws.cpu().numpy()[:, :, :n]).max() < 2e-3
    assert np.array_equal(outs[0].stat("is_accepted").cpu().numpy()[:, :, :n],
                          outs[1].stat("is_accepted").cpu().numpy()[:, :, :n])
    assert np.allclose(outs[0].stat("step_size").cpu().numpy()[:, :, :n],
                       outs[1].stat("step_size").cpu().numpy()[:, :, :n], rtol=1e-3)
    return outs, pk, prob, init


@pytest.mark.parametrize("method", [0, 1])
@pytest.mark.parametrize("kind", ["levels3", "factorial2x2", "conditions4", "factorial2x2x2", "groups3_d2", "conditions5",
                                  "factorial2x2x2x2"])
def test_multi_group_kernel_follows_generic_kernel(kind, method):
    """The lane-per-cell-type kernel for up to 16 covariate groups and D <= 4 (hmc_mg.cuh) against the generic kernel and
    the oracle's log-density, on designs with 3 to 16 groups."""
    K = 9
    xs, ys = zip(*[_multi_group_dataset(80 + b, kind, N=32, K=K, zero_frac=0.1 * (b == 1)) for b in range(3)])
    outs, pk, prob, init = _mg_vs_generic(xs, ys, [K - 1, 0, 4], method=method)
    assert 2 < pk.Umax <= 16
    # chunked resume is bit-identical to the single launch
    smp = prob.make_sampler(method, 12, 0, num_adapt=12, seed=4, warps_per_chain=1)
    many = prob.sample(init, smp, chunk=5)
    assert many.launches > 1 and np.array_equal(many.draws.cpu().numpy(), outs[1].draws.cpu().numpy())


def test_multi_group_kernel_ragged_batch_and_cell_type_range():
    """One batch whose datasets have 4, 3, 2 and 1 covariate groups (the batch-wide layout is padded to 4), with
    pseudocounts, and the largest cell-type count the kernel takes (K = 31)."""
    N = 16
    xs, ys = [], []
    for b, kind in enumerate(["conditions4", "conditions4", "conditions4", "conditions4"]):
        x, y = _multi_group_dataset(90 + b, kind, N=N, K=31, zero_frac=0.15 if b == 2 else 0.0)
        if b == 1:
            x[x[:, 2] == 1] = 0.0              # three groups
        if b == 2:
            x[:, 1:] = 0.0                     # two groups
        if b == 3:
            x[:] = 0.0                         # a single group
        xs.append(x)
        ys.append(y)
    outs, pk, _, _ = _mg_vs_generic(xs, ys, [30, 0, 7, 12], C=3, steps=8)
    assert sorted(pk.U.tolist()) == [1, 2, 3, 4] and pk.NOT >= 1


@pytest.mark.parametrize("K,N,kind", [(2, 64, "factorial2x2x2"), (3, 90, "levels3"), (12, 120, "conditions4")])
def test_multi_group_kernel_long_groups_and_few_cell_types(K, N, kind):
    """Several items per pair (groups of 8 to 30 rows) and the smallest cell-type counts on the multi-group kernel."""
    xs, ys = zip(*[_multi_group_dataset(130 + b, kind, N=N, K=K) for b in range(2)])
    outs, pk, _, _ = _mg_vs_generic(xs, ys, [K - 1, 0], steps=8)
    assert pk.NT >= 2


@pytest.mark.parametrize("N,grouped_expected", [(1000, 1), (2000, 0)])
def test_long_case_control_dataset_runs_on_the_generic_kernels(N, grouped_expected):
    """Datasets too long for the lane-per-cell-type tile (N = 1000) or for the pair/item layout altogether (N = 2000)
    still run — on the generic kernels — and trace the oracle's log-density."""
    from oracle import np_oracle as o
    from sccoda_b200._pack import pack
    from sccoda_b200.engine import DeviceProblem
    K, C, steps = 20, 2, 4
    x, y = synthetic_dataset(7, N=N, K=K)
    pk = pack(x, y, K - 1, dtype=np.float32)
    assert pk.grouped == grouped_expected
    prob = DeviceProblem(pk, chains=C)
    init = _init(np.random.RandomState(1), 1, K, C).reshape(1, C, -1) * 0.3
    init[..., 0] = 1.0
    out = prob.sample(init, prob.make_sampler(0, steps, 0, num_adapt=steps, seed=2, step_size=0.002))
    assert out.kernel_kind == 1
    d = out.draws.double().cpu().numpy()
    lp = out.stat("target_log_prob").double().cpu().numpy()
    assert np.isfinite(d).all() and np.isfinite(lp).all()
    m = o.prepare(x, y, K - 1)
    for c in range(C):
        ref = np.array([o.logp(d[0, c, i], m) for i in range(steps)])
        assert np.abs(lp[0, c] - ref).max() < 5e-6 * np.abs(ref).max() + 4e-3, np.abs(lp[0, c] - ref).max()


def test_multi_group_kernel_frozen_spike_slab():
    """SimpleModel semantics (sigma_d and ind_raw frozen) on the multi-group kernel."""
    from sccoda_b200._pack import pack
    from sccoda_b200.engine import DeviceProblem
    K = 6
    x, y = _multi_group_dataset(5, "factorial2x2", N=20, K=K)
    prob = DeviceProblem(pack(x[None], y[None], np.array([K - 1]), dtype=np.float32), chains=4)
    init = _init(np.random.RandomState(2), 2, K, 4).reshape(1, 4, -1) * 0.3
    init[..., :2] = 1.0
    init[..., 2 + 2 * (K - 1):2 + 4 * (K - 1)] = 0.7
    outs = [prob.sample(init, prob.make_sampler(0, 30, 0, num_adapt=30, seed=9, warps_per_chain=1, generic_only=g,
                                                freeze_spike_slab=True)) for g in (True, False)]
    assert outs[1].kernel_kind == 5
    d = outs[1].draws.cpu().numpy()
    assert np.all(d[..., :2] == 1.0) and np.all(d[..., 2 + 2 * (K - 1):2 + 4 * (K - 1)] == np.float32(0.7))
    assert np.abs(outs[0].draws.cpu().numpy()[:, :, :5] - d[:, :, :5]).max() < 2e-3


def test_fp32_statistical_parity_multi_group():
    """Posterior means and accept rates of the multi-group kernel against the generic kernel over many chains."""
    from sccoda_b200._pack import pack
    from sccoda_b200.engine import DeviceProblem
    K, C = 6, 256
    x, y = _multi_group_dataset(12, "conditions4", N=16, K=K)
    prob = DeviceProblem(pack(x[None], y[None], np.array([K - 1]), dtype=np.float32), chains=C)
    init = np.zeros((1, C, 3 + 6 * (K - 1) + K), np.float32)
    init[..., :3] = 1.0
    init[..., 3:3 + 3 * (K - 1)] = np.random.RandomState(0).randn(C, 3 * (K - 1))
    res = []
    for g in (True, False):
        out = prob.sample(init, prob.make_sampler(0, 1500, 500, seed=21, warps_per_chain=1, generic_only=g))
        assert out.kernel_kind == (1 if g else 5)
        d = out.draws.double().cpu().numpy()[0, :, 500:]
        acc = out.stat("is_accepted").double().cpu().numpy()[0].mean()
        res.append((d.mean(axis=1), acc))
    (m0, a0), (m1, a1) = res
    assert abs(a0 - a1) < 0.02
    se = np.sqrt(m0.var(axis=0) / C + m1.var(axis=0) / C)
    z = np.abs(m0.mean(axis=0) - m1.mean(axis=0)) / np.maximum(se, 1e-6)
    assert z.max() < 5.0, z.max()


def test_case_control_nuts_kernel_follows_generic_kernel():
    """Same Philox streams, same tree logic: the lane-per-cell-type NUTS kernel and the generic NUTS kernel build the
    same trees for the first transitions in fp32 (later, rounding differences change a U-turn decision somewhere)."""
    from sccoda_b200._pack import pack
    from sccoda_b200.engine import DeviceProblem
    B, C, K = 2, 3, 12
    xs, ys = zip(*[synthetic_dataset(90 + b, N=14, K=K) for b in range(B)])
    ys = np.stack(ys)
    ys[1, 2, 3] = 0
    init = _init(np.random.RandomState(12), 1, K, B * C).reshape(B, C, -1) * 0.5
    init[..., 0] = 1.0
    prob = DeviceProblem(pack(np.stack(xs), ys, np.array([11, 4]), dtype=np.float32), chains=C)
    outs = [prob.sample(init, prob.make_sampler(2, 10, 0, num_adapt=10, seed=6, max_tree_depth=6, warps_per_chain=1,
                                                generic_only=g)) for g in (True, False)]
    assert outs[0].kernel_kind == 2 and outs[1].kernel_kind == 4
    d0, d1 = outs[0].draws.cpu().numpy(), outs[1].draws.cpu().numpy()
    assert np.isfinite(d1).all()
    n = 3
    for name in ("leapfrogs_taken", "is_accepted", "reach_max_depth", "diverging"):
        assert np.array_equal(outs[0].stat(name).cpu().numpy()[:, :, :n], outs[1].stat(name).cpu().numpy()[:, :, :n]), name
    assert np.abs(d0[:, :, :n] - d1[:, :, :n]).max() < 5e-3
    assert np.allclose(outs[0].stat("energy").cpu().numpy()[:, :, :n], outs[1].stat("energy").cpu().numpy()[:, :, :n], atol=5e-2)
    # chunked resume is bit-identical
    many = prob.sample(init, prob.make_sampler(2, 10, 0, num_adapt=10, seed=6, max_tree_depth=6, warps_per_chain=1), chunk=4)
    assert many.launches > 1 and np.array_equal(many.draws.cpu().numpy(), d1)
    assert np.array_equal(many.stats.cpu().numpy(), outs[1].stats.cpu().numpy(), equal_nan=True)


@pytest.mark.parametrize("kind", ["levels3", "factorial2x2", "conditions4", "factorial2x2x2", "factorial2x2x2x2"])
def test_multi_group_nuts_kernel_follows_generic_kernel(kind):
    """The lane-per-cell-type NUTS kernel for up to 16 covariate groups / D <= 4 builds the same trees as the generic NUTS
    kernel for the first transitions (same Philox streams, same tree logic), and its traced log-density is the oracle's."""
    from oracle import np_oracle as o
    from sccoda_b200._pack import pack
    from sccoda_b200.engine import DeviceProblem
    B, C, K = 2, 3, 9
    xs, ys = zip(*[_multi_group_dataset(110 + b, kind, N=32, K=K, zero_frac=0.1 * (b == 1)) for b in range(B)])
    D = xs[0].shape[1]
    refs = [K - 1, 3]
    init = _init(np.random.RandomState(12), D, K, B * C).reshape(B, C, -1) * 0.5
    init[..., :D] = 1.0
    prob = DeviceProblem(pack(np.stack(xs), np.stack(ys), np.array(refs), dtype=np.float32), chains=C)
    outs = [prob.sample(init, prob.make_sampler(2, 10, 0, num_adapt=10, seed=6, max_tree_depth=6, warps_per_chain=1,
                                                generic_only=g)) for g in (True, False)]
    assert outs[0].kernel_kind == 2 and outs[1].kernel_kind == 6
    d0, d1 = outs[0].draws.cpu().numpy(), outs[1].draws.cpu().numpy()
    assert np.isfinite(d1).all()
    n = 3
    for name in ("leapfrogs_taken", "is_accepted", "reach_max_depth", "diverging"):
        assert np.array_equal(outs[0].stat(name).cpu().numpy()[:, :, :n], outs[1].stat(name).cpu().numpy()[:, :, :n]), name
    assert np.abs(d0[:, :, :n] - d1[:, :, :n]).max() < 5e-3
    assert np.allclose(outs[0].stat("energy").cpu().numpy()[:, :, :n], outs[1].stat("energy").cpu().numpy()[:, :, :n], atol=5e-2)
    lp = outs[1].stat("target_log_prob").double().cpu().numpy()
    for b in range(B):
        m = o.prepare(xs[b], ys[b], int(refs[b]))
        for c in range(C):
            ref = np.array([o.logp(d1[b, c, i].astype(np.float64), m) for i in range(10)])
            assert np.abs(lp[b, c] - ref).max() < 5e-6 * np.abs(ref).max() + 4e-3, (b, c)
    # chunked resume is bit-identical
    many = prob.sample(init, prob.make_sampler(2, 10, 0, num_adapt=10, seed=6, max_tree_depth=6, warps_per_chain=1), chunk=4)
    assert many.launches > 1 and np.array_equal(many.draws.cpu().numpy(), d1)
    assert np.array_equal(many.stats.cpu().numpy(), outs[1].stats.cpu().numpy(), equal_nan=True)


@pytest.mark.parametrize("variant", ["generic", "case_control"])
def test_fp32_statistical_parity_nuts(variant):
    from sccoda_b200.engine import DeviceProblem
    x, y, _ = haber_salm()
    C, nr, nb = 32, 400, 300
    init = _init(np.random.RandomState(8), 1, 8, C)
    d_ref, s_ref = _oracle_run(x, y, 5, init, 2, nr, nb, seed=31, chain_id0=0, max_tree_depth=8)
    prob = DeviceProblem.from_arrays(x, y, 5, chains=C, dtype=np.float32)
    cc = variant == "case_control"
    out = prob.sample(init[None], prob.make_sampler(2, nr, nb, seed=32, max_tree_depth=8, warps_per_chain=1 if cc else 0,
                                                    generic_only=not cc))
    assert out.kernel_kind == (4 if cc else 2)
    d = out.draws.double().cpu().numpy()
    assert np.isfinite(d).all()
    lf, lf_r = out.stat("leapfrogs_taken").double().cpu().numpy()[0].mean(1), s_ref["leapfrogs_taken"][0].mean(1)
    assert abs(np.log(lf.mean() / lf_r.mean())) < 0.35, (lf.mean(), lf_r.mean())
    a = np.exp(np.minimum(out.stat("log_accept_ratio").double().cpu().numpy()[0], 0)).mean()
    a_r = np.exp(np.minimum(s_ref["log_accept_ratio"][0], 0)).mean()
    assert abs(a - a_r) < 0.08, (a, a_r)
    am, ar = d[0, :, :, -8:].mean(1), d_ref[0, :, :, -8:].mean(1)
    se = np.hypot(am.std(0, ddof=1), ar.std(0, ddof=1)) / np.sqrt(C)
    assert np.all(np.abs(am.mean(0) - ar.mean(0)) < 5 * se + 0.05), (am.mean(0), ar.mean(0))


def _config3_data():
    """BASELINE config 3: N = 200 samples, K = 50 cell types, D = 4 binary covariates (SURVEY §8d recipe)."""
    rs = np.random.RandomState(1234)
    N, K, D = 200, 50, 4
    x = rs.randint(0, 2, size=(N, D)).astype(float)
    bt = rs.uniform(-3, 3, size=K)
    wt = np.zeros((D, K))
    for dd in range(D):
        wt[dd, rs.choice(K - 1, 3, replace=False)] = rs.choice([0.3, 0.5, 1.0], 3)
    logits = x @ wt + bt + np.sqrt(0.05) * rs.randn(N, K)
    p = np.exp(logits - logits.max(1, keepdims=True))
    p /= p.sum(1, keepdims=True)
    y = np.stack([rs.multinomial(5000, p[n]) for n in range(N)]).astype(float)
    return x, y


def test_config3_shape_nuts_fp64_trajectory_and_fp32_statistics():
    """BASELINE config 3 (N = 200, K = 50, D = 4, sample_nuts, 16 warps per chain): the fp64 build walks the oracle's
    first NUTS transitions, and the fp32 production build agrees with the oracle statistically (leapfrogs per draw, accept
    statistic, step size, intercept means) on short chains with dual averaging (the long version of the comparison is
    tools/nuts_cfg3_parity.py -> profiles/r2_nuts_cfg3_parity.json; the oracle needs minutes for it)."""
    from sccoda_b200.engine import DeviceProblem
    x, y = _config3_data()
    K, D, C = 50, 4, 4
    init = _init(np.random.RandomState(4), D, K, C)
    init[:, -K:] = np.log(y.mean(0) / y.mean(0).sum() * 50.0) + 0.1 * init[:, -K:]    # intercepts near the data: no long transient
    # fp64: identical Philox streams, trees of up to 2^6 leaves
    d_ref, s_ref = _oracle_run(x, y, K - 1, init[:2], 2, 4, 0, num_adapt=4, seed=5, chain_id0=1, max_tree_depth=6)
    prob64 = DeviceProblem.from_arrays(x, y, K - 1, chains=2, dtype=np.float64)
    out = prob64.sample(init[None, :2], prob64.make_sampler(2, 4, 0, num_adapt=4, seed=5, chain_id0=1, max_tree_depth=6))
    assert out.kernel_kind == 2 and out.warps_per_chain == 16
    assert np.array_equal(out.stat("leapfrogs_taken").cpu().numpy(), s_ref["leapfrogs_taken"])
    assert np.abs(out.draws.cpu().numpy() - d_ref).max() < 1e-6
    # fp32: the config's sampler (depth 10, dual averaging), independent streams
    nr, nb = 20, 60
    d_ref, s_ref = _oracle_run(x, y, K - 1, init, 2, nr, nb, seed=41, chain_id0=0)
    prob = DeviceProblem.from_arrays(x, y, K - 1, chains=C, dtype=np.float32)
    out = prob.sample(init[None], prob.make_sampler(2, nr, nb, seed=42))
    assert out.kernel_kind == 2 and out.warps_per_chain == 16
    d = out.draws.double().cpu().numpy()
    assert np.isfinite(d).all()
    lf, lf_r = out.stat("leapfrogs_taken").double().cpu().numpy()[0].mean(), s_ref["leapfrogs_taken"][0].mean()
    assert abs(np.log(lf / lf_r)) < 0.35, (lf, lf_r)
    a = np.exp(np.minimum(out.stat("log_accept_ratio").double().cpu().numpy()[0], 0)).mean()
    a_r = np.exp(np.minimum(s_ref["log_accept_ratio"][0], 0)).mean()
    assert abs(a - a_r) < 0.08, (a, a_r)
    eps, eps_r = out.stat("step_size").double().cpu().numpy()[0, :, -1].mean(), s_ref["step_size"][0, :, -1].mean()
    assert abs(np.log(eps / eps_r)) < 0.35, (eps, eps_r)
    am, ar = d[0, :, :, -K:].mean(1), d_ref[0, :, :, -K:].mean(1)                  # [C,K] intercept means per chain
    se = np.hypot(am.std(0, ddof=1), ar.std(0, ddof=1)) / np.sqrt(C)
    assert np.all(np.abs(am.mean(0) - ar.mean(0)) < 5 * se + 0.05), np.abs(am.mean(0) - ar.mean(0)).max()
    lp, lp_r = out.stat("target_log_prob").double().cpu().numpy()[0], s_ref["target_log_prob"][0]
    assert abs(lp.mean() - lp_r.mean()) < 5 * np.hypot(lp.std(), lp_r.std()) / np.sqrt(C) + 25.0, (lp.mean(), lp_r.mean())


def test_summary_kernel_matches_numpy():
    """K4 vs the numpy restatement of get_y_hat / complete_beta_df arithmetic on the same draws."""
    from oracle import np_oracle as o
    from sccoda_b200.engine import DeviceProblem
    x, y = synthetic_dataset(3, N=10, K=6, D=2)
    C, nr, nb, skip = 3, 500, 200, 150
    init = _init(np.random.RandomState(9), 2, 6, C)
    prob = DeviceProblem.from_arrays(x, y, 4, chains=C, dtype=np.float32)
    out = prob.sample(init[None], prob.make_sampler(0, nr, nb, seed=5))
    s = prob.split_summary(prob.summarize(out.draws, out.stats, skip).cpu().numpy())
    d = out.draws.double().cpu().numpy()[0]
    m = o.prepare(x, y, 4)
    for c in range(C):
        used = d[c, skip:]
        assert np.allclose(s["mean"][0, c], used.mean(0), rtol=1e-6, atol=1e-7)
        assert np.allclose(s["m2"][0, c], ((used - used.mean(0)) ** 2).sum(0), rtol=1e-5, atol=1e-6)
        der, _ = o.derived_draws(used, m)
        beta = der["beta"].reshape(len(used), -1)
        inc = (np.abs(beta) > 1e-3)
        assert np.abs(s["inc_count"][0, c] - inc.sum(0)).max() <= 2          # |beta| ~ 1e-3 ties in fp32
        assert np.allclose(s["beta_mean"][0, c], beta.mean(0), atol=1e-5)
        assert np.allclose(s["inc_sum"][0, c], (beta * inc).sum(0), atol=5e-3)
        assert s["n_used"][0, c] == nr - skip
        assert s["n_accepted"][0, c] == out.stat("is_accepted")[0, c].sum().item()


def test_host_buffer_entry_point_matches_device_path():
    """sccoda_sample_host (host pointers, copies inside) == the device-pointer path, bit for bit."""
    from sccoda_b200._pack import pack
    from sccoda_b200.engine import DeviceProblem, sample_host
    x, y, _ = haber_salm()
    pk = pack(x, y, 5, dtype=np.float32)
    init = _init(np.random.RandomState(3), 1, 8, 4).astype(np.float32)
    prob = DeviceProblem(pk, chains=4)
    smp = prob.make_sampler(1, 200, 100, seed=17)
    out = prob.sample(init[None], smp)
    summ_dev = prob.summarize(out.draws, out.stats, 50).cpu().numpy()
    draws, stats, summ, ms = sample_host(pk, 4, init[None], prob.make_sampler(1, 200, 100, seed=17),
                                         want_draws=True, want_summary=True, summary_skip=50)
    assert ms > 0
    assert np.array_equal(draws, out.draws.cpu().numpy())
    assert np.array_equal(stats, out.stats.cpu().numpy(), equal_nan=True)
    assert np.array_equal(summ, summ_dev)


@pytest.mark.parametrize("B,C", [(640, 4), (900, 3), (2500, 1), (828, 5), (2072, 4), (400, 4)])   # two full waves; 11 chains per SM
def test_wide_launch_is_bit_identical_to_one_dataset_per_cta(B, C):
    """Launches with more than 8 chains per SM take the wide geometry of the case/control HMC kernel (one CTA per SM
    with a balanced share of the chains, staging every dataset they belong to, barrier every 16 transitions); chains do
    not depend on the launch geometry: draws, statistics and the resume state are those of the classic launch."""
    from sccoda_b200._pack import pack
    from sccoda_b200.engine import DeviceProblem
    from sccoda_b200.util.data_generation import generate_sweep
    x, y, _ = generate_sweep(B, K=20, n_per_group=10, n_total=5000, seed0=77)
    y[3, 2, 5] = 0
    y[B - 1, 19, 0] = 0
    refs = np.arange(B) % 20
    prob = DeviceProblem(pack(x, y, refs, dtype=np.float32), chains=C)
    rs = np.random.RandomState(B)
    init = np.zeros((B, C, prob.P), dtype=np.float32)
    init[..., 0] = 1.0
    init[..., 1:20] = rs.randn(B, C, 19)
    init[..., -20:] = rs.randn(B, C, 20)
    smp_w = prob.make_sampler(1, 40, 20, seed=5)
    smp_c = prob.make_sampler(1, 40, 20, seed=5, chains_per_cta=min(C, 4))
    assert prob.plan(smp_w)[1] > 8 and prob.plan(smp_c)[1] == min(C, 4)
    ow, oc = prob.sample(init, smp_w), prob.sample(init, smp_c)
    assert ow.kernel_kind == 3 and oc.kernel_kind == 3
    assert torch_equal(ow.draws, oc.draws) and torch_equal(ow.stats, oc.stats)
    P = prob.P
    assert torch_equal(ow.carry[..., :P + 5], oc.carry[..., :P + 5])       # slots P+5, P+6 are timing diagnostics
    assert 0.2 < ow.stat("is_accepted").mean().item() < 1.0
    if C == 4:
        # chunked launches resume through the carry (each chunk is its own wide launch, barriers restart at the chunk's
        # first transition) and the SimpleModel variant (frozen sigma_d / ind_raw) takes the same geometry
        och = prob.sample(init, prob.make_sampler(1, 40, 20, seed=5), chunk=23)
        assert och.launches == 3 and torch_equal(och.draws, oc.draws) and torch_equal(och.stats, oc.stats)
        fw = prob.sample(init, prob.make_sampler(0, 30, 10, seed=9, freeze_spike_slab=True))
        fc = prob.sample(init, prob.make_sampler(0, 30, 10, seed=9, freeze_spike_slab=True, chains_per_cta=4))
        assert torch_equal(fw.draws, fc.draws) and bool((fw.draws[..., 0] == 1.0).all())


@pytest.mark.parametrize("conditions,K,B,C", [(3, 20, 520, 4), (4, 12, 700, 3), (3, 40, 440, 4)])
def test_wide_launch_of_the_multi_group_kernel_is_bit_identical(conditions, K, B, C):
    """The same geometry for the multi-group HMC kernels (one CTA per SM of 4 x the resident CTAs' warps): one-hot coded
    conditions, K <= 31 and two cell types per lane; bit-identical to one dataset per CTA."""
    from sccoda_b200._pack import pack
    from sccoda_b200.engine import DeviceProblem
    from sccoda_b200.util.data_generation import generate_sweep
    N = 5 * conditions
    _, y, _ = generate_sweep(B, K=K, n_per_group=(N + 1) // 2, n_total=5000, seed0=99)
    y = y[:, :N]
    y[1, 2, 3] = 0
    lv = np.arange(N) % conditions
    x = np.broadcast_to((lv[:, None] == np.arange(1, conditions)[None, :]).astype(float), (B, N, conditions - 1)).copy()
    D = conditions - 1
    prob = DeviceProblem(pack(x, y, K - 1, dtype=np.float32), chains=C)
    rs = np.random.RandomState(B)
    init = np.zeros((B, C, prob.P), dtype=np.float32)
    init[..., :D] = 1.0
    init[..., D:D + D * (K - 1)] = rs.randn(B, C, D * (K - 1))
    init[..., -K:] = rs.randn(B, C, K)
    smp_w = prob.make_sampler(0, 30, 20, seed=6)
    smp_c = prob.make_sampler(0, 30, 20, seed=6, chains_per_cta=min(C, 4))
    assert prob.plan(smp_w)[1] > 4 and prob.plan(smp_c)[1] == min(C, 4)
    ow, oc = prob.sample(init, smp_w), prob.sample(init, smp_c)
    assert ow.kernel_kind == 5 and oc.kernel_kind == 5
    assert torch_equal(ow.draws, oc.draws) and torch_equal(ow.stats, oc.stats)
    assert torch_equal(ow.carry[..., :prob.P + 5], oc.carry[..., :prob.P + 5])


def torch_equal(a, b):
    import torch
    return bool(torch.equal(torch.nan_to_num(a, nan=-7.0), torch.nan_to_num(b, nan=-7.0)))


def test_raw_host_entry_point_packs_and_matches_device_path():
    """sccoda_sample_raw_host (raw x / y host arrays: packer + copies + sampler + summaries inside one native call) ==
    packing in Python and sampling through the device-pointer path, bit for bit."""
    from sccoda_b200._pack import pack
    from sccoda_b200.engine import DeviceProblem, sample_raw_host
    xs, ys = zip(*[synthetic_dataset(60 + b, N=20, K=20) for b in range(3)])
    xs, ys = np.stack(xs), np.stack(ys)
    ys[1, 3, 2] = 0                                                          # pseudocount inside the native packer
    refs = np.array([19, 0, 7], dtype=np.int32)
    init = np.stack([_init(np.random.RandomState(3 + b), 1, 20, 4) for b in range(3)]).astype(np.float32)
    prob = DeviceProblem(pack(xs, ys, refs, dtype=np.float32), chains=4)
    out = prob.sample(init, prob.make_sampler(0, 150, 50, seed=23))
    summ_dev = prob.summarize(out.draws, out.stats, 50).cpu().numpy()
    draws, stats, summ, tm = sample_raw_host(xs, ys, refs, 4, init, prob.make_sampler(0, 150, 50, seed=23),
                                             want_draws=True, want_summary=True, summary_skip=50)
    assert tm["kernel_ms"] > 0 and tm["pack_ms"] > 0 and tm["total_ms"] >= tm["pack_ms"]
    assert np.array_equal(draws, out.draws.cpu().numpy())
    assert np.array_equal(stats, out.stats.cpu().numpy(), equal_nan=True)
    assert np.array_equal(summ, summ_dev)


def test_diverging_flag_and_negative_sigma_rejection():
    """A huge step size throws proposals into sigma_d < 0 / overflow: they must be rejected, flagged and finite."""
    from sccoda_b200.engine import DeviceProblem
    x, y, _ = haber_salm()
    init = _init(np.random.RandomState(5), 1, 8, 8)
    prob = DeviceProblem.from_arrays(x, y, 5, chains=8, dtype=np.float32)
    out = prob.sample(init[None], prob.make_sampler(0, 200, 0, num_adapt=0, step_size=0.8, seed=1))
    assert np.isfinite(out.draws.cpu().numpy()).all()
    lar = out.stat("log_accept_ratio").cpu().numpy()
    div = out.stat("diverging").cpu().numpy()
    assert np.array_equal(div == 1, lar < -1000)                          # scCODA_model.py:299
    acc = out.stat("is_accepted").cpu().numpy()
    assert acc[lar == -np.inf].sum() == 0


@pytest.mark.parametrize("seed", range(6))
def test_random_shapes_trace_the_oracle_log_density(seed):
    """Randomised shapes across the dispatch table (case/control, multi-group and generic kernels; HMC, HMC-DA and NUTS;
    binary, one-hot, continuous and constant designs; counts from sparse to large): whatever kernel the planner picks,
    the traced log-density is the oracle's log-density at the traced state."""
    from oracle import np_oracle as o
    from sccoda_b200._pack import pack
    from sccoda_b200.engine import DeviceProblem
    rs = np.random.RandomState(1000 + seed)
    seen = set()
    for trial in range(5):
        N = int(rs.choice([3, 6, 12, 25, 60]))
        K = int(rs.choice([2, 3, 5, 8, 20, 31, 32, 40]))
        D = int(rs.choice([1, 2, 3, 4, 5]))
        mode = int(rs.randint(4))
        if mode == 0:
            x = rs.randint(0, 2, size=(N, D)).astype(float)
        elif mode == 1:
            x = rs.randn(N, D)
        elif mode == 2:
            x = (rs.randint(0, D + 1, size=N)[:, None] == np.arange(1, D + 1)[None, :]).astype(float)
        else:
            x = np.zeros((N, D))
            x[N // 2:, 0] = 1.0
        y = rs.poisson(rs.choice([0.7, 4.0, 30.0, 400.0]), size=(N, K)).astype(float)
        y[0, 0] += 1.0
        ref = int(rs.randint(K))
        C, steps = 2, 5
        method = int(rs.randint(3))
        pk = pack(x, y, ref, dtype=np.float32)
        prob = DeviceProblem(pk, chains=C)
        init = _init(rs, D, K, C).reshape(1, C, -1) * 0.3
        init[..., :D] = 1.0
        out = prob.sample(init, prob.make_sampler(method, steps, 0, num_adapt=steps, seed=seed, step_size=0.005,
                                                  max_tree_depth=4))
        seen.add(out.kernel_kind)
        d = out.draws.double().cpu().numpy()
        lp = out.stat("target_log_prob").double().cpu().numpy()
        assert np.isfinite(d).all() and np.isfinite(lp).all(), (N, K, D, mode, method, out.kernel_kind)
        m = o.prepare(x, y, ref)
        for c in range(C):
            want = np.array([o.logp(d[0, c, i], m) for i in range(steps)])
            err = np.abs(lp[0, c] - want).max()
            assert err < 5e-6 * np.abs(want).max() + 4e-3, (N, K, D, mode, method, out.kernel_kind, err)
    assert len(seen) >= 2
