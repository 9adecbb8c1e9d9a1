"""CPU tests: the C-ABI library loads and exports every symbol include/sccoda_b200.h declares; struct layouts and
argument validation behave (no kernel is launched here)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, synthetic_dataset
from sccoda_b200 import _native as nat
from sccoda_b200._pack import pack


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "sccoda_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sccoda_[a-z_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = nat.lib()
    names = _declared_functions()
    assert set(names) == set(nat.EXPORTS), (names, nat.EXPORTS)
    for n in names:
        assert hasattr(lib, n), n
    assert lib.sccoda_version() >= 100


def test_header_constants_match_binding():
    src = open(os.path.join(ROOT, "include", "sccoda_b200.h")).read()
    consts = dict(re.findall(r"#define\s+(SCCODA_[A-Z0-9_]+)\s+(\d+)", src))
    assert int(consts["SCCODA_NSTAT"]) == nat.NSTAT == len(nat.STAT_NAMES)
    assert int(consts["SCCODA_NCARRY"]) == nat.NCARRY
    assert int(consts["SCCODA_NSUM_PARAM"]) == nat.NSUM_PARAM
    assert int(consts["SCCODA_NSUM_BETA"]) == nat.NSUM_BETA
    assert int(consts["SCCODA_NSUM_SCALAR"]) == nat.NSUM_SCALAR
    for i, name in enumerate(nat.STAT_NAMES):
        key = "SCCODA_ST_" + name.upper()
        assert int(consts[key]) == i, key


def _problem(pk, C):
    return nat.Problem(C=C, **nat.problem_fields(pk, lambda name: getattr(pk, name).ctypes.data))


def _sampler(**kw):
    d = dict(method=0, num_results=100, num_burnin=50, num_adapt=40, num_leapfrog=10, max_tree_depth=10, step_begin=0,
             step_end=0, step_size=0.01, target_accept=0.75, seed=1, chain_id0=0, warps_per_chain=0, store_draws=1,
             chains_per_cta=0, generic_only=0, freeze_spike_slab=0, reject_conc_above=0.0)
    d.update(kw)
    return nat.Sampler(**d)


def _plan(prob, smp):
    w, c, g, s = (ctypes.c_int32() for _ in range(4))
    rc = nat.lib().sccoda_launch_plan(ctypes.byref(prob), ctypes.byref(smp), *(ctypes.byref(v) for v in (w, c, g, s)))
    return rc, w.value, c.value, g.value, s.value


def test_launch_plan_sweep_config():
    """1024 datasets x 4 chains, K=N=20: one warp per chain; the sweep fills the device, so the planner takes the wide
    launch — one CTA per SM with a balanced share of the 4096 chains (28 warps at most), staging the up to 8 datasets
    they belong to.  A shard of 128 datasets (strong scaling on 8 GPUs) keeps one dataset per CTA."""
    x, y = synthetic_dataset(0)
    pk = pack(np.broadcast_to(x, (1024,) + x.shape), np.broadcast_to(y, (1024,) + y.shape), 19)
    rc, w, cpc, grid, smem = _plan(_problem(pk, 4), _sampler())
    assert rc == 0 and w == 1 and cpc == 28 and grid == 148 and 100 * 1024 < smem < 227 * 1024
    rc, w, cpc, grid, smem = _plan(_problem(pk, 4), _sampler(chains_per_cta=4))          # an explicit geometry is kept
    assert rc == 0 and w == 1 and cpc == 4 and grid == 1024 and 0 < smem < 48 * 1024
    pk = pack(np.broadcast_to(x, (128,) + x.shape), np.broadcast_to(y, (128,) + y.shape), 19)
    rc, w, cpc, grid, smem = _plan(_problem(pk, 4), _sampler())
    assert rc == 0 and w == 1 and cpc == 4 and grid == 128 and 0 < smem < 48 * 1024
    pk = pack(np.broadcast_to(x, (512,) + x.shape), np.broadcast_to(y, (512,) + y.shape), 19)      # two GPUs: 14 chains per SM
    rc, w, cpc, grid, smem = _plan(_problem(pk, 4), _sampler())
    assert rc == 0 and w == 1 and cpc == 14 and grid == 148


def test_launch_plan_scales_warps_for_few_chains_and_big_models():
    x, y = synthetic_dataset(0, N=200, K=50, D=4)
    pk = pack(x, y, 49)
    rc, w, cpc, grid, smem = _plan(_problem(pk, 64), _sampler(method=2))
    assert rc == 0 and w == 16 and cpc == 1 and grid == 64 and smem <= 227 * 1024
    # a single chain of a small case/control problem: the lane-per-cell-type kernel (one warp) ...
    x, y = synthetic_dataset(0, N=10, K=5)
    pk = pack(x, y, 4)
    rc, w, cpc, grid, smem = _plan(_problem(pk, 1), _sampler())
    assert rc == 0 and w == 1 and grid == 1 and pk.grouped == 1
    kind = ctypes.c_int32()
    assert nat.lib().sccoda_kernel_kind(ctypes.byref(_problem(pk, 1)), ctypes.byref(_sampler()), ctypes.byref(kind)) == 0
    assert kind.value == 3
    # ... NUTS has its lane-per-cell-type kernel too ...
    rc, w, *_ = _plan(_problem(pk, 1), _sampler(method=2))
    assert rc == 0 and w == 1
    assert nat.lib().sccoda_kernel_kind(ctypes.byref(_problem(pk, 1)), ctypes.byref(_sampler(method=2)), ctypes.byref(kind)) == 0
    assert kind.value == 4
    # ... but several warps per chain when the sampler asks for the generic kernels
    rc, w, *_ = _plan(_problem(pk, 1), _sampler(generic_only=1))
    assert rc == 0 and w == 2


def test_kernel_dispatch_table():
    """Which kernel the library picks (sccoda_kernel_kind): case/control, multi-group (one or two cell types per
    lane), generic."""
    def kinds(x, y, ref, C=4096):
        pk = pack(x, y, ref)
        out = []
        for method in (0, 2):
            k = ctypes.c_int32()
            assert nat.lib().sccoda_kernel_kind(ctypes.byref(_problem(pk, C)), ctypes.byref(_sampler(method=method)),
                                                ctypes.byref(k)) == 0
            out.append(k.value)
        return tuple(out)

    rs = np.random.RandomState(0)
    x, y = synthetic_dataset(0, N=12, K=20)
    assert kinds(x, y, 19) == (3, 4)                                        # case/control, K <= 31
    x, y = synthetic_dataset(0, N=12, K=40)
    assert kinds(x, y, 39) == (5, 6)                                        # case/control, 32 <= K <= 62: two columns per lane
    x, y = synthetic_dataset(0, N=12, K=63)
    assert kinds(x, y, 62) == (1, 2)                                        # K > 62: generic
    x4 = ((np.arange(32)[:, None] >> np.arange(4)[None, :]) & 1).astype(float)
    _, y = synthetic_dataset(0, N=32, K=20)
    assert kinds(x4, y, 19) == (5, 6)                                       # 2^4 factorial: D = 4, 16 groups
    _, y = synthetic_dataset(0, N=32, K=40)
    assert kinds(x4, y, 39) == (1, 2)                                       # D = 4 with K > 31: generic
    assert kinds(x4[:, :3], y, 39) == (5, 6)                                # D = 3 with K <= 62
    x5 = ((np.arange(64)[:, None] >> np.arange(5)[None, :]) & 1).astype(float)
    _, y = synthetic_dataset(0, N=64, K=20)
    assert kinds(x5, y, 19) == (1, 2)                                       # 32 groups / D = 5: generic
    _, y = synthetic_dataset(0, N=12, K=20)
    assert kinds(rs.randn(12, 1), y, 19) == (5, 6)                          # continuous covariate, 12 rows = 12 groups
    _, y = synthetic_dataset(0, N=40, K=20)
    assert kinds(rs.randn(40, 1), y, 19) == (1, 2)                          # ... 40 groups: generic
    _, y = synthetic_dataset(0, N=24, K=20)
    xc = rs.randn(24, 1)
    assert kinds(xc, y, 19) == (5, 6)                                       # 24 groups, a full device: one warp per chain
    assert kinds(xc, y, 19, C=4) == (1, 2)                                  # ... a handful of chains: several warps per chain


def test_launch_plan_falls_back_when_the_specialised_tile_does_not_fit():
    """Long case/control datasets: the lane-per-cell-type kernels stage the whole element list (16 B per count) in
    shared memory; when that does not fit, the planner takes the generic kernels instead of failing, and the packer
    leaves the pair/item layout when even those cannot hold it."""
    kind = ctypes.c_int32()
    x, y = synthetic_dataset(0, N=200, K=20)
    pk = pack(x, y, 19)
    assert pk.grouped == 1
    assert nat.lib().sccoda_kernel_kind(ctypes.byref(_problem(pk, 4096)), ctypes.byref(_sampler()), ctypes.byref(kind)) == 0
    assert kind.value == 3
    x, y = synthetic_dataset(0, N=1000, K=20)
    pk = pack(x, y, 19)
    assert pk.grouped == 1
    for method, generic_kind in ((0, 1), (2, 2)):
        rc, w, cpc, grid, smem = _plan(_problem(pk, 4096), _sampler(method=method))
        assert rc == 0 and w > 1 and smem <= 227 * 1024
        assert nat.lib().sccoda_kernel_kind(ctypes.byref(_problem(pk, 4096)), ctypes.byref(_sampler(method=method)),
                                            ctypes.byref(kind)) == 0
        assert kind.value == generic_kind
    x, y = synthetic_dataset(0, N=2000, K=20)
    assert pack(x, y, 19, grouped=True).grouped == 1
    assert _plan(_problem(pack(x, y, 19, grouped=True), 4), _sampler(method=2))[0] == -9
    pk = pack(x, y, 19)                                  # automatic choice: element layout
    assert pk.grouped == 0
    rc, w, cpc, grid, smem = _plan(_problem(pk, 4), _sampler(method=2))
    assert rc == 0 and smem <= 227 * 1024


@pytest.mark.parametrize("kw,code", [(dict(method=7), -11), (dict(num_results=0), -12), (dict(num_leapfrog=0), -13),
                                      (dict(method=2, max_tree_depth=0), -14), (dict(step_size=0.0), -15),
                                      (dict(step_begin=10, step_end=5), -16), (dict(warps_per_chain=3), -8)])
def test_sampler_validation(kw, code):
    x, y = synthetic_dataset(0)
    pk = pack(x, y, 19)
    rc, *_ = _plan(_problem(pk, 4), _sampler(**kw))
    assert rc == code
    assert len(nat.lib().sccoda_last_error()) > 0


def test_problem_validation():
    x, y = synthetic_dataset(0)
    pk = pack(x, y, 19)
    p = _problem(pk, 4)
    p.K = 1
    assert _plan(p, _sampler())[0] == -3
    p = _problem(pk, 0)
    assert _plan(p, _sampler())[0] == -2
    p = _problem(pk, 4)
    p.ey = None
    assert _plan(p, _sampler())[0] == -7
    assert pk.grouped == 1
    p = _problem(pk, 4)
    p.pcoef = None
    assert _plan(p, _sampler())[0] == -7
    p = _problem(pk, 4)
    p.NImax = 7                       # must be even
    assert _plan(p, _sampler())[0] == -6


def test_problem_struct_matches_header():
    """Field order of the ctypes mirror == field order of `struct sccoda_problem` in the header."""
    src = open(os.path.join(ROOT, "include", "sccoda_b200.h")).read()
    body = re.search(r"typedef struct sccoda_problem \{(.*?)\} sccoda_problem;", src, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if decl:
            names += [n.strip().lstrip("*") for n in decl.split(",")]
            names[-len(decl.split(",")):] = [re.split(r"[\s\*]+", n.strip())[-1] for n in decl.split(",")]
    assert names == [f[0] for f in nat.Problem._fields_], names


def test_check_raises_runtime_error():
    with pytest.raises(RuntimeError, match="failed"):
        nat.check(-3, "unit-test")


def test_product_never_touches_the_oracle_and_fails_loudly_without_the_library(tmp_path, monkeypatch):
    """The oracle is test infrastructure: nothing under sccoda_b200/ may import, load or execute anything under oracle/
    (only tests/, __graft_entry__.smoke() and bench.py's CPU legs do), and without the native library every entry point
    raises instead of falling back."""
    import os
    import re
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|libdm_oracle|oracle[/\\]", re.M)
    for dirpath, _, files in os.walk(os.path.join(root, "sccoda_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".sh")):
                text = open(os.path.join(dirpath, f), encoding="utf-8", errors="replace").read()
                assert not pat.search(text), os.path.join(dirpath, f)
    code = ("import numpy as np\n"
            "from sccoda_b200._pack import pack\n"
            "try:\n"
            "    pack(np.zeros((4, 1)), np.ones((4, 3)), 0)\n"
            "except RuntimeError as e:\n"
            "    assert 'no CPU fallback' in str(e).replace('There is no', 'no'), e\n"
            "    print('raised')\n")
    env = dict(os.environ, SCCODA_B200_LIB=str(tmp_path / "missing.so"), PYTHONPATH=root)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=120)
    assert out.returncode == 0 and "raised" in out.stdout, out.stderr[-500:]
