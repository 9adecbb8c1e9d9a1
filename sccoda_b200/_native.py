"""
ctypes binding of the C ABI in include/sccoda_b200.h (libsccoda_b200.so, built in-tree by csrc/build.sh).

There is NO fallback: if the shared library is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SCCODA_B200_LIB overrides the library path (development builds, e.g. ablation variants)
LIB_PATH = os.environ.get("SCCODA_B200_LIB") or os.path.join(_HERE, "libsccoda_b200.so")

F32, F64 = 0, 1
METHOD_HMC, METHOD_HMC_DA, METHOD_NUTS = 0, 1, 2
NSTAT = 8
STAT_NAMES = ("target_log_prob", "log_accept_ratio", "is_accepted", "step_size", "diverging",
              "leapfrogs_taken", "energy", "reach_max_depth")
NCARRY = 8
NSUM_PARAM, NSUM_BETA, NSUM_SCALAR = 2, 4, 6

EXPORTS = ("sccoda_version", "sccoda_last_error", "sccoda_logp_grad", "sccoda_logp_grad_as", "sccoda_sample", "sccoda_summarize", "sccoda_predict",
           "sccoda_launch_plan", "sccoda_kernel_kind", "sccoda_sample_host", "sccoda_release_workspace", "sccoda_measure_peaks",
           "sccoda_pack", "sccoda_packed_problem", "sccoda_packed_row_order", "sccoda_packed_y", "sccoda_packed_logp_const",
           "sccoda_packed_free", "sccoda_sample_raw_host")


class Problem(ctypes.Structure):
    _fields_ = [("B", ctypes.c_int32), ("C", ctypes.c_int32), ("N", ctypes.c_int32), ("K", ctypes.c_int32),
                ("D", ctypes.c_int32), ("Umax", ctypes.c_int32), ("dtype", ctypes.c_int32),
                ("nbig_min", ctypes.c_int32),
                ("grouped", ctypes.c_int32), ("NImax", ctypes.c_int32), ("NOmax", ctypes.c_int32),
                ("NGmax", ctypes.c_int32), ("NF", ctypes.c_int32), ("NT", ctypes.c_int32), ("NOT", ctypes.c_int32),
                ("reserved0", ctypes.c_int32),
                ("ey", ctypes.c_void_p), ("eycst", ctypes.c_void_p), ("eidx", ctypes.c_void_p),
                ("colperm", ctypes.c_void_p), ("nbig", ctypes.c_void_p),
                ("ntot", ctypes.c_void_p), ("xu", ctypes.c_void_p),
                ("grp", ctypes.c_void_p), ("rstart", ctypes.c_void_p), ("U", ctypes.c_void_p),
                ("ref", ctypes.c_void_p), ("lconst", ctypes.c_void_p),
                ("NI", ctypes.c_void_p), ("iy", ctypes.c_void_p), ("ipair", ctypes.c_void_p),
                ("ipos", ctypes.c_void_p), ("inreal", ctypes.c_void_p), ("pcoef", ctypes.c_void_p), ("NG", ctypes.c_void_p),
                ("oy", ctypes.c_void_p), ("opair", ctypes.c_void_p), ("opos", ctypes.c_void_p),
                ("odesc", ctypes.c_void_p)]


def problem_fields(pk, ptr):
    """Keyword arguments of Problem(...) for a PackedProblem; ptr(array_or_tensor) -> address."""
    names = ("ey", "eycst", "eidx", "colperm", "nbig", "ntot", "xu", "grp", "rstart", "U", "ref", "lconst",
             "NI", "iy", "ipair", "ipos", "inreal", "pcoef", "NG", "oy", "opair", "opos", "odesc")
    kw = dict(B=pk.B, N=pk.N, K=pk.K, D=pk.D, Umax=pk.Umax, dtype=F32 if pk.dtype.itemsize == 4 else F64,
              nbig_min=pk.nbig_min, grouped=pk.grouped, NImax=pk.NImax, NOmax=pk.NOmax, NGmax=pk.NGmax, NF=pk.NF, NT=pk.NT, NOT=pk.NOT)
    kw.update({n: ptr(n) for n in names})
    return kw


class Sampler(ctypes.Structure):
    _fields_ = [("method", ctypes.c_int32), ("num_results", ctypes.c_int32), ("num_burnin", ctypes.c_int32),
                ("num_adapt", ctypes.c_int32), ("num_leapfrog", ctypes.c_int32), ("max_tree_depth", ctypes.c_int32),
                ("step_begin", ctypes.c_int32), ("step_end", ctypes.c_int32),
                ("step_size", ctypes.c_double), ("target_accept", ctypes.c_double),
                ("seed", ctypes.c_uint64), ("chain_id0", ctypes.c_uint32), ("warps_per_chain", ctypes.c_int32),
                ("store_draws", ctypes.c_int32), ("chains_per_cta", ctypes.c_int32),
                ("generic_only", ctypes.c_int32), ("freeze_spike_slab", ctypes.c_int32),
                ("reject_conc_above", ctypes.c_double)]


_lib = None


def lib() -> ctypes.CDLL:
    """Load libsccoda_b200.so; raise loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"sccoda_b200: native library {LIB_PATH} is missing. Build it with "
                f"`python -c 'import __graft_entry__ as g; g.build()'` or sccoda_b200/csrc/build.sh. "
                f"There is no CPU fallback.")
        _lib = ctypes.CDLL(LIB_PATH)
        vp, i32 = ctypes.c_void_p, ctypes.c_int32
        pp, sp = ctypes.POINTER(Problem), ctypes.POINTER(Sampler)
        _lib.sccoda_version.restype = ctypes.c_int
        _lib.sccoda_last_error.restype = ctypes.c_char_p
        _lib.sccoda_logp_grad.argtypes = [pp, vp, vp, vp, vp]
        _lib.sccoda_logp_grad_as.argtypes = [pp, i32, ctypes.c_double, vp, vp, vp, vp]
        _lib.sccoda_sample.argtypes = [pp, sp, vp, vp, vp, vp, vp]
        _lib.sccoda_summarize.argtypes = [pp, i32, i32, vp, vp, vp, vp]
        _lib.sccoda_predict.argtypes = [pp, vp, i32, i32, i32, vp, vp, vp, vp]
        _lib.sccoda_launch_plan.argtypes = [pp, sp] + [ctypes.POINTER(i32)] * 4
        _lib.sccoda_kernel_kind.argtypes = [pp, sp, ctypes.POINTER(i32)]
        _lib.sccoda_sample_host.argtypes = [pp, sp, vp, vp, vp, vp, i32, i32, ctypes.POINTER(ctypes.c_float)]
        _lib.sccoda_measure_peaks.argtypes = [ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float), vp]
        _lib.sccoda_pack.argtypes = [i32, i32, i32, i32, vp, vp, vp, i32, i32, i32, i32, ctypes.POINTER(vp)]
        _lib.sccoda_packed_problem.argtypes = [vp, i32, pp]
        for name in ("sccoda_packed_row_order", "sccoda_packed_y", "sccoda_packed_logp_const"):
            getattr(_lib, name).argtypes = [vp]
            getattr(_lib, name).restype = vp
        _lib.sccoda_packed_free.argtypes = [vp]
        _lib.sccoda_packed_free.restype = None
        _lib.sccoda_sample_raw_host.argtypes = [i32, i32, i32, i32, i32, vp, vp, vp, i32, sp, vp, vp, vp, vp, i32, i32,
                                                ctypes.POINTER(ctypes.c_float)]
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().sccoda_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"sccoda_b200 native call {what} failed (code {rc}): {msg}")
