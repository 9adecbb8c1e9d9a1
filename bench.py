#!/usr/bin/env python
"""
bench.py — headline benchmark of the B200 MCMC engine (contract: see the task statement / DESIGN.md §Measurement).

Workload (BASELINE.json configs[4], the one the metric is quoted on; it fits one GPU):
    1024 independent synthetic case-control datasets (K=20 cell types, N=20 samples, D=1) x 4 chains,
    sample_hmc semantics: 5,000 burn-in + 20,000 traced HMC steps of 10 leapfrogs (250,000 gradient evaluations
    per chain, 1.024e9 per 1024-dataset sweep).
One "step" = one full sweep of that workload.  With N GPUs the 1024 datasets are sharded over the ranks (STRONG scaling,
the configuration as BASELINE.json writes it: 128 datasets x 4 chains = 512 chains per GPU at N = 8); dataset ids, seeds
and Philox streams are global, so the result does not depend on N.  No data-path collective: the only exchange is the
gather of the per-chain summaries (inside the timed region when N > 1).  For N > 1 the weak-scaling figure of round 1
(1024 datasets per GPU) is measured as well and reported under "weak".

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "leapfrog_grad_evals_per_sec"
UNIT = "grad-evals/s"


# ------------------------------------------------------------------------------------------------------------
def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--datasets", type=int, default=1024, help="datasets of the sweep, TOTAL over all GPUs (strong scaling)")
    ap.add_argument("--weak", action="store_true", help="--datasets per GPU instead (weak scaling, the round-1 headline)")
    ap.add_argument("--weak-extra-steps", type=int, default=2,
                    help="N > 1: also time this many sweeps of the weak workload (--datasets per GPU), reported under 'weak'")
    ap.add_argument("--chains", type=int, default=4)
    ap.add_argument("--num-results", type=int, default=20000)
    ap.add_argument("--num-burnin", type=int, default=5000)
    ap.add_argument("--leapfrog", type=int, default=10)
    ap.add_argument("--K", type=int, default=20)
    ap.add_argument("--n-per-group", type=int, default=10)
    ap.add_argument("--warps-per-chain", type=int, default=0)
    ap.add_argument("--method", default="hmc", choices=["hmc", "hmc_da", "nuts"])
    ap.add_argument("--ess-datasets", type=int, default=-1, help="datasets per rank in the ESS/s figure (-1 = all, 0 = skip)")
    ap.add_argument("--cpu-steps", type=int, default=1500, help="HMC steps per CPU-baseline chain")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def flops_per_grad(N, K, D):
    """Canonical algorithmic work of one gradient evaluation: 64 NK + 48 N + 4 NKD flop."""
    return 64 * N * K + 48 * N + 4 * N * K * D


def flops_per_logp(N, K):
    """Canonical work of one value evaluation: 28 flop per lgamma, 2NK + 2N lgammas."""
    return 28 * (2 * N * K + 2 * N)


def workload_name(a):
    where = "per GPU" if a.weak else "in total, sharded over the GPUs"
    return (f"simulation sweep: {a.datasets} datasets {where} x {a.chains} chains, K={a.K} N={2 * a.n_per_group} D=1, "
            f"sample_{a.method}({a.num_results},{a.num_burnin}) L={a.leapfrog}")


# ------------------------------------------------------------------------------------------------------------
# CPU legs (the oracle port; never on the product path)
# ------------------------------------------------------------------------------------------------------------
def cpu_numpy_port(a, n_proc, steps):
    """numpy/scipy restatement, one chain per process, n_proc processes (SURVEY §8d CPU-baseline definition)."""
    import multiprocessing as mp
    from oracle import cpu_baseline
    ctx = mp.get_context("spawn")
    tasks = [(i, a.K, a.n_per_group, steps, a.leapfrog) for i in range(n_proc)]
    # one chain per process, one BLAS/OpenMP thread per process: n_proc cores busy, no oversubscription
    saved = {k: os.environ.get(k) for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS")}
    for k in saved:
        os.environ[k] = "1"
    try:
        pool = ctx.Pool(n_proc)
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    with pool:
        pool.map(cpu_baseline.warm, range(n_proc))               # imports + first-call overheads, untimed
        t0 = time.perf_counter()
        res = pool.map(cpu_baseline.run_numpy_chain, tasks)
        wall = time.perf_counter() - t0
    ge = sum(r[0] for r in res)
    # the n_proc chains run concurrently, one per core: aggregate rate = sum of the per-process rates
    # (each worker times only its own chain, not process start-up / imports / data generation)
    rate = sum(r[0] / r[1] for r in res)
    return rate, ge, wall


def cpu_c_port(a, n_threads, steps):
    from oracle import cpu_baseline
    return cpu_baseline.run_c_port(a.K, a.n_per_group, steps, a.leapfrog, n_threads)


def cpu_ess_per_s(a, n_threads, num_results=4000, num_burnin=1000):
    """ESS/s of the CPU path: one sample_hmc-style chain per host thread (C + OpenMP port of the oracle, the fastest CPU
    statement of the reference's sampler), burn-in : results = 1 : 4 as in the benchmark workload; bulk ESS of every
    chain after the reference's second burn-in, minimum / median over the coordinates, summed over the chains, divided
    by the wall time of the whole run.  Returns a dict."""
    from oracle import cpu_baseline
    return cpu_baseline.run_c_port_ess(a.K, a.n_per_group, num_results, num_burnin, a.leapfrog, n_threads)


def run_reference(a):
    """--impl reference: the reference's CPU implementation of the path.  TensorFlow(-Probability) is not
    installable here (no wheel, no network), so this is the oracle PORT (numpy/scipy fp64 restatement of the
    TFP sampler), with every host core, on a bounded sample of the same workload per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    steps_per_sample = 250                       # HMC steps per chain per bench step (bounded sample)
    vals = []
    for i in range(a.warmup + a.steps):
        v, ge, wall = cpu_numpy_port(a, cores, steps_per_sample)
        if i >= a.warmup:
            vals.append((v, wall))
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([w for _, w in vals]) * 1e3)
    c_val, _, _ = cpu_c_port(a, cores, 2000)
    ess = cpu_ess_per_s(a, cores)
    sample = (f"{cores} processes x 1 chain x {steps_per_sample} HMC steps x {a.leapfrog} leapfrogs on datasets "
              f"0..{cores - 1} of the sweep (numpy/scipy fp64 port of the TFP sampler)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak" if a.weak else "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a), "note": "TensorFlow absent: oracle port (numpy/scipy) times the path"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "c_port_value": c_val, "published_tf_cpu_anchor": 3.3e3, "ess": ess},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [s.strip() for s in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "power_w_max": float(max(power)) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


PACKED_DEVICE_FIELDS = ("ey", "eycst", "eidx", "colperm", "nbig", "ntot", "xu", "grp", "rstart", "U", "ref", "lconst",
                        "NI", "iy", "ipair", "ipos", "inreal", "pcoef", "NG", "oy", "opair", "opos", "odesc")


def traffic_from_profile(kernel_kind, a, world=1):
    """dram__bytes_read + dram__bytes_write of the sampler launch from the committed `ncu --set full` capture
    (profiles/r2_traffic.json, else r1_traffic.json); only valid for the default workload on one GPU and the kernel it was
    taken on."""
    p = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if not os.path.exists(p):
        p = os.path.join(ROOT, "profiles", "r1_traffic.json")
    if world != 1 or kernel_kind != 3 or (a.num_results, a.num_burnin, a.datasets, a.chains, a.K, a.n_per_group) != (20000, 5000, 1024, 4, 20, 10) or not os.path.exists(p):
        return None
    with open(p) as f:
        t = json.load(f)
    return {"bytes": t["dram_bytes_read"] + t["dram_bytes_write"], "unit": "B per launch", "source": t["source"],
            "ncu_pipes_pct": {"fma_cycles_active": t.get("sm__pipe_fma_cycles_active_pct"),
                              "fma_inst_executed": t.get("sm__inst_executed_pipe_fma_pct"),
                              "xu_inst_executed": t.get("sm__inst_executed_pipe_xu_pct"),
                              "issue_active": t.get("smsp__issue_active_pct")}}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


def pinned_copy(packed):
    """Move the packed host arrays into pinned memory (the e2e path copies from pinned host buffers)."""
    import dataclasses
    import torch
    keep = []

    def pin(arr):
        arr = np.ascontiguousarray(arr)
        signed = {np.dtype(np.uint16): np.int16, np.dtype(np.uint32): np.int32}.get(arr.dtype)
        t = torch.from_numpy(arr.view(signed) if signed else arr).pin_memory()
        keep.append(t)
        return t.numpy().view(arr.dtype)
    fields = {f.name: getattr(packed, f.name) for f in dataclasses.fields(packed)}
    for k in PACKED_DEVICE_FIELDS:
        fields[k] = pin(fields[k])
    return type(packed)(**fields), keep


# ------------------------------------------------------------------------------------------------------------
def run_ours(a):
    import ctypes

    import torch
    from sccoda_b200 import _native as nat
    from sccoda_b200 import scheduler as sch
    from sccoda_b200._pack import pack
    from sccoda_b200.engine import DeviceProblem, sample_raw_host
    from sccoda_b200.util.data_generation import generate_sweep

    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1 and a.gpus > 1:
        emit({"error": f"--gpus {a.gpus} needs torchrun with {a.gpus} ranks"})
        sys.exit(2)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    # CPU baseline first (rank 0, N=1 only), before CUDA is initialised in this process
    cpu_baseline = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cores = os.cpu_count() or 1
        v, ge, wall = cpu_numpy_port(a, cores, a.cpu_steps)
        c_val, _, _ = cpu_c_port(a, cores, 4 * a.cpu_steps)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": (f"{cores} processes x 1 chain x {a.cpu_steps} HMC steps x {a.leapfrog} leapfrogs on "
                                   f"datasets 0..{cores - 1} of the sweep, numpy/scipy fp64 port of the TFP sampler "
                                   f"({ge} grad-evals in {wall:.1f}s); TensorFlow is not installable here"),
                        "c_port_value": c_val, "c_port_note": "same algorithm, plain C + OpenMP, all cores",
                        "published_tf_cpu_anchor": 3.3e3,
                        "published_tf_cpu_anchor_note": "reference tutorials: 3.0-3.6k grad-evals/s, 1 chain (BASELINE.md)",
                        "ess": cpu_ess_per_s(a, cores)}

    comm = sch.Comm.from_env()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    method = {"hmc": nat.METHOD_HMC, "hmc_da": nat.METHOD_HMC_DA, "nuts": nat.METHOD_NUTS}[a.method]
    C, K, N, D = a.chains, a.K, 2 * a.n_per_group, 1
    total_steps = a.num_burnin + a.num_results
    skip = min(a.num_burnin, a.num_results)
    seed = 20261019

    class Shard:
        """This rank's datasets [lo, hi) of a sweep, resident on the device, and one bench step over them."""

        def __init__(self, lo, hi, n_total_datasets):
            self.lo, self.hi, self.B, self.n_total = lo, hi, hi - lo, n_total_datasets
            self.x, self.y, _ = generate_sweep(self.B, K=K, n_per_group=a.n_per_group, n_total=5000, seed0=1234, dataset0=lo)
            self.packed = pack(self.x, self.y, K - 1, dtype=np.float32)
            self.init = sch.initial_states(self.B, C, D, K, seed=1234, dataset0=lo).astype(np.float32)
            self.prob = DeviceProblem(self.packed, chains=C)
            self.P = self.packed.P
            self.draws = torch.empty((self.B, C, a.num_results, self.P), dtype=torch.float32, device=dev)
            self.stats = torch.empty((self.B, C, a.num_results, nat.NSTAT), dtype=torch.float32, device=dev)
            self.init_dev = torch.from_numpy(self.init).to(dev)
            self.smp = self.prob.make_sampler(method, a.num_results, a.num_burnin, None, a.leapfrog, 10, 0.01, 0.75, seed=seed,
                                              chain_id0=lo * C, warps_per_chain=a.warps_per_chain)
            self.phase_events = []

        def step(self):
            """One sweep: launch 1 = the persistent sampler (K2), launch 2 = the summaries (K4), then — N > 1 — the only
            exchange, the gather of the per-chain summaries.  No host synchronisation inside."""
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            out = self.prob.sample(self.init_dev, self.smp, draws=self.draws, stats=self.stats, sync=False)
            ev[0].record()
            summ = self.prob.summarize(out.draws, out.stats, skip)
            ev[1].record()
            if world > 1:
                comm.gather_blocks(summ.reshape(self.B, -1), self.n_total)
            ev[2].record()
            self.phase_events.append((out.events[0], out.events[1], ev[0], ev[1], ev[2]))
            return out, summ

        def phases_ms(self):
            """mean device time per step of (sampler kernel, summary kernel, gather) on this rank"""
            k = np.mean([e[0].elapsed_time(e[1]) for e in self.phase_events])
            sm = np.mean([e[2].elapsed_time(e[3]) for e in self.phase_events])
            g = np.mean([e[3].elapsed_time(e[4]) for e in self.phase_events])
            return float(k), float(sm), float(g)

        def grad_evals(self, out):
            if method == nat.METHOD_NUTS:
                return float(out.stat("leapfrogs_taken").double().sum().item()) * total_steps / a.num_results
            return float(self.B) * C * total_steps * a.leapfrog

        def timed(self, warmup, steps, clocks=None):
            for _ in range(warmup):
                out, summ = self.step()
            torch.cuda.synchronize()
            self.phase_events.clear()
            comm.barrier()
            torch.cuda.synchronize()
            if clocks is not None:
                clocks.start()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                out, summ = self.step()
            e1.record()
            comm.barrier()
            torch.cuda.synchronize()
            info = clocks.stop() if clocks is not None else None
            ms_per_step = comm.max_float(e0.elapsed_time(e1)) / steps
            ge_all = comm.sum_float(self.grad_evals(out))
            return out, summ, ms_per_step, ge_all, info

    if a.weak:
        lo, hi, n_all = rank * a.datasets, (rank + 1) * a.datasets, a.datasets * world
    else:
        lo, hi = sch.shard_range(a.datasets, rank, world)
        n_all = a.datasets
    if hi <= lo:
        emit({"error": f"rank {rank} has no dataset: --datasets {a.datasets} < world size {world}"})
        sys.exit(2)
    sh = Shard(lo, hi, n_all)
    B, P, prob = sh.B, sh.P, sh.prob

    # roofline denominators measured on this device
    f1, f2 = ctypes.c_float(), ctypes.c_float()
    nat.check(nat.lib().sccoda_measure_peaks(ctypes.byref(f1), ctypes.byref(f2), None), "sccoda_measure_peaks")
    fp32_peak, mufu_peak = float(f1.value), float(f2.value)

    out, summ, ms_per_step, ge_all, clock_info = sh.timed(a.warmup, a.steps, ClockSampler(local_rank))
    value = ge_all / (ms_per_step * 1e-3)
    k_ms, s_ms, g_ms = sh.phases_ms()
    # every rank's phase times, so that a slow GPU or a slow exchange can be named from the record
    per_rank = None
    if world > 1:
        t = torch.tensor([[k_ms, s_ms, g_ms]], dtype=torch.float64, device=dev)
        allr = comm.gather_blocks(t, world).cpu().numpy()
        per_rank = {"kernel_ms": [round(float(v), 3) for v in allr[:, 0]], "summary_ms": [round(float(v), 3) for v in allr[:, 1]],
                    "gather_ms": [round(float(v), 3) for v in allr[:, 2]],
                    "note": "mean device time per step and rank (CUDA events); gather_ms includes waiting for the slowest rank"}

    # ---- roofline of the dominant kernel (the persistent sampler), per launch, on this rank
    ge_rank = sh.grad_evals(out)
    flops_launch = (B * C * total_steps * (a.leapfrog * flops_per_grad(N, K, D) + flops_per_logp(N, K))
                    if method != nat.METHOD_NUTS else ge_rank * (flops_per_grad(N, K, D) + flops_per_logp(N, K)))
    achieved_tf = flops_launch / (k_ms * 1e-3) / 1e12
    peaks, peak_kind = measured_peaks()
    draw_bytes = float(B) * C * a.num_results * (P + nat.NSTAT) * 4
    traffic = traffic_from_profile(out.kernel_kind, a, world)
    plan = prob.plan(sh.smp)
    wide = out.kernel_kind == 3 and plan[1] > C                 # one CTA per SM with a share of all chains (hmc_cc.cuh)
    n_warps = plan[2] * plan[1] * plan[0]
    cc_name = (("hmc_cc_wide_kernel" if plan[1] > 16 else "hmc_cc_wide16_kernel") if wide else
               "hmc_cc_lat_kernel" if n_warps <= 148 * 12 else
               "hmc_cc_lat16_kernel" if n_warps <= 148 * 16 else "hmc_cc_kernel")       # kernels_cc.cu: launch_hmc_cc
    roofline = {"bound": "fp32", "achieved": achieved_tf, "peak": fp32_peak, "unit": "TFLOP/s",
                "frac": achieved_tf / fp32_peak if fp32_peak > 0 else None,
                "traffic": traffic["bytes"] if traffic else None, "traffic_detail": traffic,
                "peak_source": "FFMA micro-kernel measured in this run on this device (sccoda_measure_peaks); "
                               "MEASURED_PEAKS.json has no fp32 entry",
                "mufu_peak_tops": mufu_peak,
                "kernel": {1: "hmc_kernel<float,W,grouped>", 3: cc_name,
                           5: "hmc_mg_kernel<D>"}.get(out.kernel_kind, "hmc_kernel"), "kernel_ms": k_ms,
                "summary_kernel_ms": s_ms, "algorithmic_flop_per_launch": flops_launch,
                "flop_model": "per grad-eval 64NK+48N+4NKD, per value 28(2NK+2N) (SURVEY 8d)",
                "hbm": {"achieved_gbs": draw_bytes / (k_ms * 1e-3) / 1e9, "peak_gbs": peaks.get("hbm_gbs"),
                        "peak_source": peak_kind, "algorithmic_bytes_per_launch": draw_bytes,
                        "note": "draw write-out only; the path is FP32/SFU-bound, not HBM-bound"}}

    # ---- ESS/s over ALL datasets of every rank (post-processing of the device-resident draws, outside the timed region)
    ess_info = None
    n_ess = B if a.ess_datasets < 0 else min(a.ess_datasets, B)
    if n_ess > 0 and C > 1:
        try:
            t0 = time.perf_counter()
            cfg = sch.SweepConfig(method=method, num_results=a.num_results, num_burnin=a.num_burnin, num_leapfrog=a.leapfrog,
                                  seed=seed, chains=C, dtype=np.float32, diag_datasets=n_ess)
            res = sch.shard_results(sch.SweepShard(lo=lo, hi=hi, problem=prob, out=out, summary=summ), cfg)
            ess = res["ess_bulk"][:n_ess]                                       # [n, coord] device
            rh = res["rhat_rank"][:n_ess]
            ess_min, ess_med = torch.nan_to_num(ess, nan=float("inf")).min(dim=1).values, ess.nanmedian(dim=1).values
            sums = torch.stack([ess_min.sum(), ess_med.sum(), torch.tensor(float(n_ess), dtype=torch.float64, device=dev)])
            if world > 1:
                import torch.distributed as dist
                dist.all_reduce(sums)
                rh_max = torch.tensor([float(torch.nan_to_num(rh, nan=0.0).max())], dtype=torch.float64, device=dev)
                dist.all_reduce(rh_max, op=dist.ReduceOp.MAX)
                rhat_max = float(rh_max.item())
            else:
                rhat_max = float(torch.nan_to_num(rh, nan=0.0).max())
            torch.cuda.synchronize()
            n_used = float(sums[2].item())
            scale = float(n_all) / n_used                                     # 1.0 unless --ess-datasets limits the sample
            ess_info = {"datasets_used_all_ranks": int(n_used), "datasets_total": int(n_all),
                        "coords": "alpha, beta(non-ref), target_log_prob",
                        "ess_min_coord_mean_over_datasets": float(sums[0].item() / n_used),
                        "ess_median_coord_mean_over_datasets": float(sums[1].item() / n_used),
                        "ess_min_coord_median_this_rank": float(ess_min.median().item()),
                        "rhat_median_this_rank": float(rh.nanmedian().item()), "rhat_max": rhat_max,
                        "ess_per_s_min_coord": float(sums[0].item() * scale / (ms_per_step * 1e-3)),
                        "ess_per_s_median_coord": float(sums[1].item() * scale / (ms_per_step * 1e-3)),
                        "compute_s": time.perf_counter() - t0,
                        "definition": "bulk ESS (rank-normalised, split chains, Geyer) over the 4 chains of a dataset after "
                                      "the reference's second burn-in, min / median over the coordinates, summed over the "
                                      "datasets of all ranks, divided by ms_per_step (the sweep samples them concurrently)"}
        except Exception as exc:  # diagnostics must never kill the bench line
            ess_info = {"error": repr(exc)}

    acc_rate = float(out.stat("is_accepted").mean().item())
    config = {"workload": workload_name(a), "datasets_this_rank": B, "grad_evals_per_step_all_gpus": ge_all,
              "warps_per_chain": plan[0], "chains_per_cta": plan[1], "grid": plan[2], "smem_bytes": plan[3],
              "l2": "inputs (a few MB) are staged once into shared memory; the draw/stat buffer written each step "
                    f"({draw_bytes / 1e9:.1f} GB per rank) is >> the 126 MB L2, so no L2 flush is needed between steps",
              "acc_rate": acc_rate}

    # ---- N > 1: the weak-scaling figure of round 1 (--datasets per GPU) beside the strong one
    weak = None
    if world > 1 and not a.weak and a.weak_extra_steps > 0:
        del out, summ
        sh.draws = sh.stats = None
        torch.cuda.empty_cache()
        shw = Shard(rank * a.datasets, (rank + 1) * a.datasets, a.datasets * world)
        outw, _, ms_w, ge_w, _ = shw.timed(1, a.weak_extra_steps)
        kw, sw, gw = shw.phases_ms()
        allw = comm.gather_blocks(torch.tensor([[kw, sw, gw]], dtype=torch.float64, device=dev), world).cpu().numpy()
        weak = {"value": ge_w / (ms_w * 1e-3), "unit": UNIT, "ms_per_step": ms_w, "datasets_per_gpu": a.datasets,
                "steps": a.weak_extra_steps, "kernel_ms_per_rank": [round(float(v), 3) for v in allw[:, 0]],
                "summary_ms_per_rank": [round(float(v), 3) for v in allw[:, 1]],
                "gather_ms_per_rank": [round(float(v), 3) for v in allw[:, 2]],
                "note": "each rank samples its own 1024 datasets (global ids rank*1024 ...): the kernel times differ with "
                        "the data, and the gather waits for the slowest rank"}
        del outw
        shw.draws = shw.stats = None
        torch.cuda.empty_cache()

    # ---- end-to-end: ONE native call from RAW host arrays (pack + H2D + sampler + summaries + D2H of the summaries)
    e2e = None
    if not a.no_e2e:
        out = summ = None
        sh.draws = sh.stats = None
        torch.cuda.empty_cache()
        x_pin = torch.from_numpy(np.ascontiguousarray(sh.x, dtype=np.float64)).pin_memory()
        y_pin = torch.from_numpy(np.ascontiguousarray(sh.y, dtype=np.float64)).pin_memory()
        init_pin = torch.from_numpy(sh.init).pin_memory()
        ref = np.full(B, K - 1, dtype=np.int32)
        h2d = sum(getattr(sh.packed, k).nbytes for k in PACKED_DEVICE_FIELDS) + init_pin.numel() * 4
        width = nat.NSUM_PARAM * P + nat.NSUM_BETA * D * K + nat.NSUM_SCALAR
        d2h = B * C * width * 8
        n_e2e = max(1, min(a.steps, 3))
        smp2 = prob.make_sampler(method, a.num_results, a.num_burnin, None, a.leapfrog, 10, 0.01, 0.75, seed=seed,
                                 chain_id0=lo * C, warps_per_chain=a.warps_per_chain)
        call = lambda: sample_raw_host(x_pin.numpy(), y_pin.numpy(), ref, C, init_pin.numpy(), smp2, want_draws=False,
                                       want_summary=True, summary_skip=skip, device=local_rank)
        call()                                                                # warm-up (allocates the cached workspace)
        comm.barrier()
        t0 = time.perf_counter()
        tms = []
        for _ in range(n_e2e):
            _, _, summ_h, tm = call()
            tms.append(tm)
        wall = comm.max_float(time.perf_counter() - t0)
        e2e = {"value": ge_all / (wall / n_e2e), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "raw_input_bytes_per_step": int(x_pin.numel() * 8 + y_pin.numel() * 8 + init_pin.numel() * 4), "steps": n_e2e,
               "pack_ms": float(np.mean([t["pack_ms"] for t in tms])), "kernel_ms": float(np.mean([t["kernel_ms"] for t in tms])),
               "total_ms": float(np.mean([t["total_ms"] for t in tms])),
               "path": "sccoda_sample_raw_host (C ABI, raw float64 x / y host arrays of this rank's datasets): C++ packer "
                       "(pseudocount, row sums, covariate groups, pair/item layout) + H2D of the packed arrays into the cached "
                       "device workspace + sampler kernel + summary kernel + D2H of the per-chain summaries, wall clock, max "
                       "over ranks; h2d/d2h bytes are the packed arrays / summaries that cross the bus"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak" if a.weak else "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "ess": ess_info,
            "gpu_launches": 2 * a.steps, "clocks": clock_info,
        }
        if per_rank is not None:
            line["per_rank"] = per_rank
        if weak is not None:
            line["weak"] = weak
        emit(line)
    comm.close()


def out_plan(prob, smp):
    return prob.plan(smp)


_REAL_STDOUT = None


def emit(line: dict) -> None:
    """Write the ONE JSON line to the real stdout (everything else — NCCL banners, warnings — goes to stderr)."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    a = parse_args()
    # keep stdout clean: C libraries (NCCL) print banners to fd 1
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
