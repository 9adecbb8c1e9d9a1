#!/usr/bin/env python
"""
bench.py — headline benchmark of the B200 MCMC engine (contract: see the task statement / DESIGN.md §Measurement).

Workload (BASELINE.json configs[4], the one the metric is quoted on; it fits one GPU):
    1024 independent synthetic case-control datasets (K=20 cell types, N=20 samples, D=1) x 4 chains,
    sample_hmc semantics: 5,000 burn-in + 20,000 traced HMC steps of 10 leapfrogs (250,000 gradient evaluations
    per chain, 1.024e9 per 1024-dataset sweep).
One "step" = one full sweep of that workload on every GPU (weak scaling: each rank runs its own 1024 datasets;
dataset ids, seeds and Philox streams are global, so ranks never repeat work).  No data-path collective: the
only exchange is the gather of per-dataset summaries (inside the timed region when N > 1).

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
    ap.add_argument("--datasets", type=int, default=1024, help="datasets per GPU")
    ap.add_argument("--chains", type=int, default=4)
    ap.add_argument("--num-results", type=int, default=20000)
    ap.add_argument("--num-burnin", type=int, default=5000)
    ap.add_argument("--leapfrog", type=int, default=10)
    ap.add_argument("--K", type=int, default=20)
    ap.add_argument("--n-per-group", type=int, default=10)
    ap.add_argument("--warps-per-chain", type=int, default=0)
    ap.add_argument("--method", default="hmc", choices=["hmc", "hmc_da", "nuts"])
    ap.add_argument("--ess-datasets", type=int, default=32, help="datasets (per rank) used for the ESS/s estimate")
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
    return (f"simulation sweep: {a.datasets} datasets/GPU x {a.chains} chains, K={a.K} N={2 * a.n_per_group} D=1, "
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
    sample = (f"{cores} processes x 1 chain x {steps_per_sample} HMC steps x {a.leapfrog} leapfrogs on datasets "
              f"0..{cores - 1} of the sweep (numpy/scipy fp64 port of the TFP sampler)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a), "note": "TensorFlow absent: oracle port (numpy/scipy) times the path"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "c_port_value": c_val, "published_tf_cpu_anchor": 3.3e3},
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


def traffic_from_profile(kernel_kind, a):
    """dram__bytes_read + dram__bytes_write of the sampler launch from the committed `ncu --set full` capture
    (profiles/r1_traffic.json); only valid for the default workload and the kernel it was taken on."""
    p = os.path.join(ROOT, "profiles", "r1_traffic.json")
    if kernel_kind != 3 or (a.num_results, a.num_burnin, a.datasets, a.chains, a.K, a.n_per_group) != (20000, 5000, 1024, 4, 20, 10) or not os.path.exists(p):
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
    import torch
    from sccoda_b200 import _native as nat
    from sccoda_b200 import scheduler as sch
    from sccoda_b200._pack import pack
    from sccoda_b200.engine import DeviceProblem, sample_host
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
                        "published_tf_cpu_anchor_note": "reference tutorials: 3.0-3.6k grad-evals/s, 1 chain (BASELINE.md)"}

    comm = sch.Comm.from_env()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    method = {"hmc": nat.METHOD_HMC, "hmc_da": nat.METHOD_HMC_DA, "nuts": nat.METHOD_NUTS}[a.method]
    B, C, K, N, D = a.datasets, a.chains, a.K, 2 * a.n_per_group, 1
    lo = rank * B                                               # weak scaling: global dataset ids [lo, lo+B)
    x, y, w_true = generate_sweep(B, K=K, n_per_group=a.n_per_group, n_total=5000, seed0=1234, dataset0=lo)
    packed = pack(x, y, K - 1, dtype=np.float32)
    init = sch.initial_states(B, C, D, K, seed=1234, dataset0=lo).astype(np.float32)
    cfg = sch.SweepConfig(method=method, num_results=a.num_results, num_burnin=a.num_burnin, num_leapfrog=a.leapfrog,
                          seed=20261019, chains=C, dtype=np.float32, warps_per_chain=a.warps_per_chain,
                          diag_datasets=0)
    prob = DeviceProblem(packed, chains=C)
    P = packed.P
    draws = torch.empty((B, C, a.num_results, P), dtype=torch.float32, device=dev)
    stats = torch.empty((B, C, a.num_results, nat.NSTAT), dtype=torch.float32, device=dev)
    init_dev = torch.from_numpy(init).to(dev)
    smp = prob.make_sampler(method, a.num_results, a.num_burnin, None, a.leapfrog, 10, 0.01, 0.75, seed=cfg.seed,
                            chain_id0=lo * C, warps_per_chain=a.warps_per_chain)
    skip = min(a.num_burnin, a.num_results)

    kernel_ms = []

    def step():
        out = prob.sample(init_dev, smp, draws=draws, stats=stats)           # launch 1: persistent sampler (K2)
        summ = prob.summarize(out.draws, out.stats, skip)                    # launch 2: summaries (K4)
        if world > 1:                                                        # the only exchange: tiny summaries
            comm.gather_blocks(summ.reshape(B, -1), B * world)
        kernel_ms.append(out.kernel_ms)
        return out, summ

    # roofline denominators measured on this device
    fp32_peak, mufu_peak = torch.zeros(1), torch.zeros(1)
    import ctypes
    f1, f2 = ctypes.c_float(), ctypes.c_float()
    nat.check(nat.lib().sccoda_measure_peaks(ctypes.byref(f1), ctypes.byref(f2), None), "sccoda_measure_peaks")
    fp32_peak, mufu_peak = float(f1.value), float(f2.value)

    for _ in range(a.warmup):
        out, summ = step()
    kernel_ms.clear()
    clocks = ClockSampler(local_rank)
    comm.barrier()
    torch.cuda.synchronize()
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        out, summ = step()
    e1.record()
    comm.barrier()
    torch.cuda.synchronize()
    clock_info = clocks.stop()
    elapsed_ms = comm.max_float(e0.elapsed_time(e1))
    ms_per_step = elapsed_ms / a.steps

    total_steps = a.num_burnin + a.num_results
    if method == nat.METHOD_NUTS:
        ge_rank = float(out.stat("leapfrogs_taken").double().sum().item()) * total_steps / a.num_results
    else:
        ge_rank = float(B) * C * total_steps * a.leapfrog
    ge_all = comm.sum_float(ge_rank)
    value = ge_all / (ms_per_step * 1e-3)

    # ---- roofline of the dominant kernel (the persistent sampler), per launch, on this rank
    k_ms = float(np.mean(kernel_ms))
    flops_launch = (B * C * total_steps * (a.leapfrog * flops_per_grad(N, K, D) + flops_per_logp(N, K))
                    if method != nat.METHOD_NUTS else ge_rank * (flops_per_grad(N, K, D) + flops_per_logp(N, K)))
    achieved_tf = flops_launch / (k_ms * 1e-3) / 1e12
    peaks, peak_kind = measured_peaks()
    draw_bytes = float(B) * C * a.num_results * (P + nat.NSTAT) * 4
    traffic = traffic_from_profile(out.kernel_kind, a)
    roofline = {"bound": "fp32", "achieved": achieved_tf, "peak": fp32_peak, "unit": "TFLOP/s",
                "frac": achieved_tf / fp32_peak if fp32_peak > 0 else None,
                "traffic": traffic["bytes"] if traffic else None, "traffic_detail": traffic,
                "peak_source": "FFMA micro-kernel measured in this run on this device (sccoda_measure_peaks); "
                               "MEASURED_PEAKS.json has no fp32 entry",
                "mufu_peak_tops": mufu_peak, "kernel": {1: "hmc_kernel<float,W,grouped>", 3: "hmc_cc_kernel", 5: "hmc_mg_kernel<D>"}.get(out.kernel_kind, "hmc_kernel"), "kernel_ms": k_ms,
                "algorithmic_flop_per_launch": flops_launch,
                "flop_model": "per grad-eval 64NK+48N+4NKD, per value 28(2NK+2N) (SURVEY 8d)",
                "hbm": {"achieved_gbs": draw_bytes / (k_ms * 1e-3) / 1e9, "peak_gbs": peaks.get("hbm_gbs"),
                        "peak_source": peak_kind, "algorithmic_bytes_per_launch": draw_bytes,
                        "note": "draw write-out only; the path is FP32/SFU-bound, not HBM-bound"}}

    # ---- ESS/s on a subset of this rank's datasets (post-processing, outside the timed region)
    ess_info = None
    if a.ess_datasets > 0:
        try:
            sh = sch.SweepShard(lo=lo, hi=lo + B, problem=prob, out=out, summary=summ)
            cfg.diag_datasets = min(a.ess_datasets, B)
            res = sch.shard_results(sh, cfg)
            ess = res["ess_bulk"][:cfg.diag_datasets].cpu().numpy()
            rh = res["rhat_rank"][:cfg.diag_datasets].cpu().numpy()
            sweep_s = k_ms * 1e-3
            per_ds_s = sweep_s  # all datasets advance concurrently: wall time of the sweep
            ess_info = {"datasets_used": int(cfg.diag_datasets), "coords": "alpha, beta(non-ref), target_log_prob",
                        "ess_min_median_over_datasets": float(np.nanmedian(np.nanmin(ess, axis=1))),
                        "ess_median_median_over_datasets": float(np.nanmedian(np.nanmedian(ess, axis=1))),
                        "rhat_median": float(np.nanmedian(rh)), "rhat_max": float(np.nanmax(rh)),
                        "ess_per_s_min_coord_all_datasets": float(np.nanmedian(np.nanmin(ess, axis=1)) * B * world / per_ds_s),
                        "ess_per_s_median_coord_all_datasets": float(np.nanmedian(np.nanmedian(ess, axis=1)) * B * world / per_ds_s),
                        "definition": "bulk ESS (rank-normalised, split chains, Geyer) over the 4 chains of a dataset "
                                      "after the reference's second burn-in; x datasets in flight / sweep time"}
        except Exception as exc:  # diagnostics must never kill the bench line
            ess_info = {"error": repr(exc)}

    acc_rate = float(out.stat("is_accepted").mean().item())

    # ---- end-to-end: the C-ABI host-buffer call (H2D of inputs, sampling, summaries, D2H of summaries)
    e2e = None
    if not a.no_e2e:
        del draws, stats, out
        torch.cuda.empty_cache()
        pk_pin, keep = pinned_copy(packed)
        init_pin = torch.from_numpy(init).pin_memory()
        h2d = sum(getattr(pk_pin, k).nbytes for k in PACKED_DEVICE_FIELDS)
        h2d += init_pin.numel() * 4
        width = nat.NSUM_PARAM * P + nat.NSUM_BETA * D * K + nat.NSUM_SCALAR
        d2h = B * C * width * 8
        smp2 = prob.make_sampler(method, a.num_results, a.num_burnin, None, a.leapfrog, 10, 0.01, 0.75, seed=cfg.seed,
                                 chain_id0=lo * C, warps_per_chain=a.warps_per_chain)
        n_e2e = max(1, min(a.steps, 3))
        sample_host(pk_pin, C, init_pin.numpy(), smp2, want_draws=False, want_summary=True, summary_skip=skip,
                    device=local_rank)                                                     # warm-up
        comm.barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            _, _, summ_h, _ = sample_host(pk_pin, C, init_pin.numpy(), smp2, want_draws=False, want_summary=True,
                                          summary_skip=skip, device=local_rank)
        wall = comm.max_float(time.perf_counter() - t0)
        e2e = {"value": ge_all / (wall / n_e2e), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "steps": n_e2e,
               "path": "sccoda_sample_host (C ABI, host pointers): cudaMalloc + H2D of packed inputs from pinned host "
                       "memory + sampler + summary kernel + D2H of per-chain summaries, wall clock"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a), "grad_evals_per_step_all_gpus": ge_all,
                       "warps_per_chain": out_plan(prob, smp)[0], "chains_per_cta": out_plan(prob, smp)[1],
                       "grid": out_plan(prob, smp)[2], "smem_bytes": out_plan(prob, smp)[3],
                       "l2": "inputs (3.5 MB) are staged once into shared memory; the 22 GB draw/stat buffer written "
                             "each step is >> the 126 MB L2, so no L2 flush is needed between steps",
                       "acc_rate": acc_rate},
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "ess": ess_info,
            "gpu_launches": 2 * a.steps, "clocks": clock_info,
        }
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
