"""
Device-side engine: owns the packed datasets in HBM (torch tensors are used only as device buffers and for
the current-stream handle) and drives the C ABI of libsccoda_b200.so.

This is the layer that replaces `CompositionalModel.sampling` -> `tfp.mcmc.sample_chain`
(sccoda/model/scCODA_model.py:115-169).  It runs B independent datasets x C chains per launch.
There is no CPU or eager-torch fallback: without a CUDA device or without the native library every entry
point raises.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _native as nat
from ._pack import PackedProblem, pack


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("sccoda_b200.engine needs a CUDA device (B200); there is no CPU fallback.")
    return torch


def _ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


@dataclass
class SampleOutput:
    draws: "object"        # torch [B,C,S,P] device
    stats: "object"        # torch [B,C,S,NSTAT] device
    carry: "object"        # torch [B,C,P+NCARRY] f64 device
    kernel_ms: float       # device time of the sampling launch(es), CUDA events
    warps_per_chain: int
    chains_per_cta: int
    grid: int
    smem_bytes: int
    launches: int
    kernel_kind: int = 0   # 1 generic HMC, 2 NUTS, 3 / 4 case/control HMC / NUTS, 5 multi-group HMC (include/sccoda_b200.h: sccoda_kernel_kind)
    events: "object" = None  # (start, end) CUDA events of the launch(es) when sample(sync=False) left kernel_ms open

    def elapsed_ms(self) -> float:
        """Device time of the sampling launch(es); synchronises on the end event when sample() did not."""
        if self.kernel_ms is None:
            self.events[1].synchronize()
            self.kernel_ms = self.events[0].elapsed_time(self.events[1])
        return self.kernel_ms

    def stat(self, name: str):
        return self.stats[..., nat.STAT_NAMES.index(name)]


class DeviceProblem:
    """B packed datasets resident on one GPU."""

    def __init__(self, packed: PackedProblem, chains: int, device: Optional[int] = None):
        torch = _torch()
        self.packed = packed
        self.C = int(chains)
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self.tdtype = torch.float32 if packed.dtype == np.float32 else torch.float64
        self.B, self.N, self.K, self.D, self.P = packed.B, packed.N, packed.K, packed.D, packed.P
        # device copies of every array of the packed layout (unsigned integer arrays travel as same-width signed views)
        signed = {np.dtype(np.uint16): np.int16, np.dtype(np.uint32): np.int32}
        self._dev = {}

        def ptr(name):
            a = np.ascontiguousarray(getattr(packed, name))
            if a.dtype in signed:
                a = a.view(signed[a.dtype])
            self._dev[name] = torch.from_numpy(a).to(self.device, non_blocking=False)
            return self._dev[name].data_ptr()

        self.c_problem = nat.Problem(C=self.C, **nat.problem_fields(packed, ptr))

    @classmethod
    def from_arrays(cls, x, y, ref, chains: int, dtype=np.float32, device: Optional[int] = None, grouped=None):
        return cls(pack(x, y, ref, dtype=dtype, grouped=grouped), chains, device)

    # ------------------------------------------------------------------ K1
    def logp_grad(self, theta, kernel_kind: Optional[int] = None, reject_conc_above: float = 0.0):
        """theta [B,C,P] (device tensor or array) -> (logp [B,C] f64, grad [B,C,P]).

        kernel_kind (as reported by `sccoda_kernel_kind` / SampleOutput.kernel_kind): evaluate through the device
        functions of that sampler kernel family — 3 / 4 the case/control kernels, 5 / 6 the multi-group kernels —
        instead of the generic ones (`sccoda_logp_grad_as`)."""
        torch = _torch()
        theta = torch.as_tensor(theta, dtype=self.tdtype, device=self.device).contiguous()
        assert tuple(theta.shape) == (self.B, self.C, self.P), theta.shape
        logp = torch.empty((self.B, self.C), dtype=torch.float64, device=self.device)
        grad = torch.empty_like(theta)
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream().cuda_stream
            if kernel_kind is None and reject_conc_above == 0.0:
                nat.check(nat.lib().sccoda_logp_grad(ctypes.byref(self.c_problem), _ptr(theta), _ptr(logp), _ptr(grad),
                                                     stream), "sccoda_logp_grad")
            else:
                nat.check(nat.lib().sccoda_logp_grad_as(ctypes.byref(self.c_problem), int(kernel_kind or 0),
                                                        float(reject_conc_above), _ptr(theta), _ptr(logp), _ptr(grad),
                                                        stream), "sccoda_logp_grad_as")
        return logp, grad

    # ------------------------------------------------------------------ K2/K3
    def make_sampler(self, method: int, num_results: int, num_burnin: int, num_adapt: Optional[int] = None,
                     num_leapfrog: int = 10, max_tree_depth: int = 10, step_size: float = 0.01,
                     target_accept: float = 0.75, seed: int = 0, chain_id0: int = 0,
                     warps_per_chain: int = 0, chains_per_cta: int = 0, generic_only: bool = False,
                     freeze_spike_slab: bool = False, reject_conc_above: float = 0.0) -> nat.Sampler:
        if num_adapt is None:
            num_adapt = int(0.8 * num_burnin)            # scCODA_model.py:284-285
        return nat.Sampler(method=method, num_results=num_results, num_burnin=num_burnin, num_adapt=num_adapt,
                           num_leapfrog=num_leapfrog, max_tree_depth=max_tree_depth, step_begin=0, step_end=0,
                           step_size=step_size, target_accept=target_accept, seed=seed, chain_id0=chain_id0,
                           warps_per_chain=warps_per_chain, store_draws=1, chains_per_cta=chains_per_cta,
                           generic_only=int(bool(generic_only)), freeze_spike_slab=int(bool(freeze_spike_slab)),
                           reject_conc_above=float(reject_conc_above))

    def plan(self, smp: nat.Sampler):
        w, c, g, s = (ctypes.c_int32() for _ in range(4))
        nat.check(nat.lib().sccoda_launch_plan(ctypes.byref(self.c_problem), ctypes.byref(smp), ctypes.byref(w),
                                               ctypes.byref(c), ctypes.byref(g), ctypes.byref(s)), "sccoda_launch_plan")
        return w.value, c.value, g.value, s.value

    def sample(self, init, smp: nat.Sampler, chunk: Optional[int] = None, progress=None,
               draws=None, stats=None, sync: bool = True) -> SampleOutput:
        """Run the whole chain (one persistent launch, or `chunk`-step launches with resume through `carry`).
        sync=False returns without waiting for the device (kernel_ms stays open until `elapsed_ms()`)."""
        torch = _torch()
        init = torch.as_tensor(init, dtype=self.tdtype, device=self.device).contiguous()
        assert tuple(init.shape) == (self.B, self.C, self.P), (init.shape, (self.B, self.C, self.P))
        S = smp.num_results
        if draws is None:
            draws = torch.empty((self.B, self.C, S, self.P), dtype=self.tdtype, device=self.device)
        if stats is None:
            stats = torch.empty((self.B, self.C, S, nat.NSTAT), dtype=self.tdtype, device=self.device)
        carry = torch.zeros((self.B, self.C, self.P + nat.NCARRY), dtype=torch.float64, device=self.device)
        total = smp.num_burnin + smp.num_results
        chunk = total if not chunk else int(chunk)
        w, cpc, grid, smem = self.plan(smp)
        kind = ctypes.c_int32()
        nat.check(nat.lib().sccoda_kernel_kind(ctypes.byref(self.c_problem), ctypes.byref(smp), ctypes.byref(kind)),
                  "sccoda_kernel_kind")
        launches = 0
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            begin = 0
            while begin < total:
                end = min(total, begin + chunk)
                smp.step_begin, smp.step_end = begin, end
                nat.check(nat.lib().sccoda_sample(ctypes.byref(self.c_problem), ctypes.byref(smp), _ptr(init),
                                                  _ptr(draws), _ptr(stats), _ptr(carry), stream.cuda_stream),
                          "sccoda_sample")
                launches += 1
                begin = end
                if progress is not None:
                    stream.synchronize()
                    progress(end, total)
            e1.record(stream)
            ms = None
            if sync:
                e1.synchronize()
                ms = e0.elapsed_time(e1)
        smp.step_begin, smp.step_end = 0, 0
        return SampleOutput(draws=draws, stats=stats, carry=carry, kernel_ms=ms, warps_per_chain=w,
                            chains_per_cta=cpc, grid=grid, smem_bytes=smem, launches=launches, kernel_kind=kind.value,
                            events=(e0, e1))

    # ------------------------------------------------------------------ K4
    def summary_width(self) -> int:
        return nat.NSUM_PARAM * self.P + nat.NSUM_BETA * self.D * self.K + nat.NSUM_SCALAR

    def summarize(self, draws, stats, skip: int):
        """Per-chain summaries [B,C,width] f64 on device (layout: include/sccoda_b200.h)."""
        torch = _torch()
        S = draws.shape[2]
        out = torch.empty((self.B, self.C, self.summary_width()), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream().cuda_stream
            nat.check(nat.lib().sccoda_summarize(ctypes.byref(self.c_problem), S, int(skip), _ptr(draws), _ptr(stats),
                                                 _ptr(out), stream), "sccoda_summarize")
        return out

    def predict(self, draws, lo: int = 0, hi: Optional[int] = None, concentration: bool = True, prediction: bool = True):
        """Posterior-predictive quantities of the draws [lo, hi) on device (K5, `sccoda_predict`): returns
        (concentration, prediction), each [B,C,hi-lo,N,K] in the problem's dtype (rows in the caller's sample order)
        or None when not requested."""
        torch = _torch()
        if getattr(self, "_row_order", None) is None:
            self._row_order = torch.as_tensor(np.ascontiguousarray(self.packed.row_order, dtype=np.int32), device=self.device)
        S = draws.shape[2]
        hi = S if hi is None else int(hi)
        shape = (self.B, self.C, hi - int(lo), self.N, self.K)
        conc = torch.empty(shape, dtype=draws.dtype, device=self.device) if concentration else None
        pred = torch.empty(shape, dtype=draws.dtype, device=self.device) if prediction else None
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream().cuda_stream
            nat.check(nat.lib().sccoda_predict(ctypes.byref(self.c_problem), _ptr(self._row_order), S, int(lo), hi, _ptr(draws),
                                               _ptr(conc) if concentration else None, _ptr(pred) if prediction else None,
                                               stream), "sccoda_predict")
        return conc, pred

    def split_summary(self, summ):
        """summary tensor/array [B,C,width] -> dict of named views."""
        P, DK = self.P, self.D * self.K
        o = 0
        out = {}
        for name, n in (("mean", P), ("m2", P), ("inc_count", DK), ("inc_sum", DK), ("beta_mean", DK),
                        ("beta_m2", DK)):
            out[name] = summ[..., o:o + n]
            o += n
        for i, name in enumerate(("n_used", "n_accepted", "sum_accept_prob", "leapfrogs", "n_diverging",
                                  "last_step_size")):
            out[name] = summ[..., o + i]
        return out


def sample_raw_host(x, y, ref, chains: int, init: np.ndarray, smp: nat.Sampler, dtype=np.float32, want_draws: bool = True,
                    want_summary: bool = False, summary_skip: int = 0, device: int = 0):
    """The whole path from RAW host arrays through `sccoda_sample_raw_host`: packing (C++), H2D, sampler, summaries,
    D2H, all inside the native call.  x [B,N,D], y [B,N,K] float64, ref [B].  Returns (draws, stats, summary, timings)
    with timings = dict(pack_ms, kernel_ms, total_ms)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    B, N, K = y.shape
    D = x.shape[2]
    assert x.shape[:2] == (B, N)
    ref = np.ascontiguousarray(np.broadcast_to(np.asarray(ref, dtype=np.int32), (B,)))
    dt = np.dtype(dtype)
    C, S, P = int(chains), smp.num_results, D + 2 * D * (K - 1) + K
    init = np.ascontiguousarray(init, dtype=dt)
    assert init.shape == (B, C, P)
    draws = np.empty((B, C, S, P), dtype=dt) if want_draws else None
    stats = np.empty((B, C, S, nat.NSTAT), dtype=dt) if want_draws else None
    width = nat.NSUM_PARAM * P + nat.NSUM_BETA * D * K + nat.NSUM_SCALAR
    summ = np.empty((B, C, width), dtype=np.float64) if want_summary else None
    tm = (ctypes.c_float * 3)()
    nat.check(nat.lib().sccoda_sample_raw_host(
        B, N, K, D, C, x.ctypes.data, y.ctypes.data, ref.ctypes.data, nat.F32 if dt.itemsize == 4 else nat.F64,
        ctypes.byref(smp), init.ctypes.data, draws.ctypes.data if draws is not None else None,
        stats.ctypes.data if stats is not None else None, summ.ctypes.data if summ is not None else None,
        int(summary_skip), int(device), tm), "sccoda_sample_raw_host")
    return draws, stats, summ, {"pack_ms": float(tm[0]), "kernel_ms": float(tm[1]), "total_ms": float(tm[2])}


def sample_host(packed: PackedProblem, chains: int, init: np.ndarray, smp: nat.Sampler, want_draws: bool = True,
                want_summary: bool = False, summary_skip: int = 0, device: int = 0):
    """End-to-end call with HOST buffers through sccoda_sample_host (copies inside the native call)."""
    dt = packed.dtype
    B, C, S, P = packed.B, int(chains), smp.num_results, packed.P
    init = np.ascontiguousarray(init, dtype=dt)
    assert init.shape == (B, C, P)
    prob = nat.Problem(C=C, **nat.problem_fields(packed, lambda name: getattr(packed, name).ctypes.data))
    draws = np.empty((B, C, S, P), dtype=dt) if want_draws else None
    stats = np.empty((B, C, S, nat.NSTAT), dtype=dt) if want_draws else None
    width = nat.NSUM_PARAM * P + nat.NSUM_BETA * packed.D * packed.K + nat.NSUM_SCALAR
    summ = np.empty((B, C, width), dtype=np.float64) if want_summary else None
    ms = ctypes.c_float(0.0)
    nat.check(nat.lib().sccoda_sample_host(
        ctypes.byref(prob), ctypes.byref(smp), init.ctypes.data,
        draws.ctypes.data if draws is not None else None, stats.ctypes.data if stats is not None else None,
        summ.ctypes.data if summ is not None else None, int(summary_skip), int(device), ctypes.byref(ms)),
        "sccoda_sample_host")
    return draws, stats, summ, float(ms.value)
