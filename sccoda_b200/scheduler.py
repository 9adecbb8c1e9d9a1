"""
S1 — multi-GPU chain / replicate scheduler.

(dataset, chain) pairs are independent units.  Datasets are sharded over ranks in contiguous,
balanced blocks so that all chains of a dataset live on one GPU (R-hat / ESS are rank-local); the Philox
streams are keyed by the GLOBAL chain id, so results do not depend on the number of GPUs.  There is no
data-path collective: the only exchange is one gather of the fixed-size per-dataset summaries / diagnostics
to rank 0 at the end (NCCL through torch.distributed; gloo in the CPU tests).

One process per GPU (torchrun); with a single process the same code runs without torch.distributed.
The reference has no counterpart: it runs one chain in one CPU process (sccoda/model/scCODA_model.py:152-161).
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Callable, Dict, Optional, Tuple

import numpy as np


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous balanced block [lo, hi) of rank; the first n_items % world ranks get one extra item."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


@dataclass
class Comm:
    """Thin wrapper over torch.distributed (or nothing, for a single process)."""
    rank: int = 0
    world: int = 1
    backend: Optional[str] = None
    device: Optional[object] = None

    @classmethod
    def from_env(cls, backend: Optional[str] = None) -> "Comm":
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if world == 1:
            return cls()
        import torch
        import torch.distributed as dist
        rank = int(os.environ["RANK"])
        local = int(os.environ.get("LOCAL_RANK", rank))
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        device = None
        if backend == "nccl":
            torch.cuda.set_device(local)
            device = torch.device("cuda", local)
        if not dist.is_initialized():
            kw = {"device_id": device} if backend == "nccl" else {}
            dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
        return cls(rank=rank, world=world, backend=backend, device=device)

    def barrier(self) -> None:
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()

    def max_float(self, v: float) -> float:
        if self.world == 1:
            return v
        import torch
        import torch.distributed as dist
        t = torch.tensor([v], dtype=torch.float64, device=self.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_float(self, v: float) -> float:
        if self.world == 1:
            return v
        import torch
        import torch.distributed as dist
        t = torch.tensor([v], dtype=torch.float64, device=self.device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def gather_blocks(self, local, n_items: int):
        """Concatenate per-rank blocks (dim 0 = this rank's shard of `n_items`) on every rank.

        Shards may differ by one item: blocks are padded to the largest shard for the fixed-size all_gather and
        trimmed afterwards."""
        if self.world == 1:
            return local
        import torch
        import torch.distributed as dist
        sizes = [shard_range(n_items, r, self.world) for r in range(self.world)]
        biggest = max(hi - lo for lo, hi in sizes)
        pad = biggest - local.shape[0]
        buf = local if pad == 0 else torch.cat([local, local.new_zeros((pad,) + tuple(local.shape[1:]))], dim=0)
        buf = buf.contiguous()
        out = [torch.empty_like(buf) for _ in range(self.world)]
        dist.all_gather(out, buf)
        return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(out, sizes)], dim=0)

    def close(self) -> None:
        if self.world > 1:
            import torch.distributed as dist
            if dist.is_initialized():
                dist.destroy_process_group()


def run_sharded(n_items: int, compute: Callable[[int, int], Dict[str, "object"]], comm: Comm) -> Dict[str, "object"]:
    """Run `compute(lo, hi)` on this rank's shard and gather every returned tensor (dim 0 = items) on all ranks."""
    lo, hi = shard_range(n_items, comm.rank, comm.world)
    local = compute(lo, hi)
    return {k: comm.gather_blocks(v, n_items) for k, v in local.items()}


# --------------------------------------------------------------------------------------------------------
# the sweep: B datasets x C chains
# --------------------------------------------------------------------------------------------------------
def initial_states(n_datasets: int, chains: int, D: int, K: int, seed: int, dataset0: int = 0) -> np.ndarray:
    """Reference initial state per chain (sccoda/model/scCODA_model.py:763-768): sigma_d = 1, b_offset ~ N(0,1),
    ind_raw = 0, alpha ~ N(0,1).  One generator per GLOBAL dataset index -> independent of the sharding."""
    P = D + 2 * D * (K - 1) + K
    out = np.zeros((n_datasets, chains, P))
    out[..., :D] = 1.0
    for i in range(n_datasets):
        rs = np.random.default_rng([seed, dataset0 + i])
        out[i, :, D:D + D * (K - 1)] = rs.standard_normal((chains, D * (K - 1)))
        out[i, :, P - K:] = rs.standard_normal((chains, K))
    return out


@dataclass
class SweepConfig:
    method: int = 0
    num_results: int = 20000
    num_burnin: int = 5000
    num_adapt: Optional[int] = None
    num_leapfrog: int = 10
    max_tree_depth: int = 10
    step_size: float = 0.01
    target_accept: float = 0.75
    seed: int = 0
    chains: int = 4
    dtype: object = np.float32
    warps_per_chain: int = 0
    est_fdr: float = 0.05
    diag_datasets: int = 0      # rank-normalised R-hat / bulk ESS for the first n local datasets (0 = none)
    hdi_prob: Optional[float] = None   # e.g. 0.94: HDIs of intercepts and effects for every dataset, on device


@dataclass
class SweepShard:
    """Everything one rank keeps after sampling its shard (device tensors)."""
    lo: int
    hi: int
    problem: object
    out: object
    summary: object
    extra: dict = field(default_factory=dict)


def sample_shard(x: np.ndarray, y: np.ndarray, ref, cfg: SweepConfig, lo: int, hi: int, draws=None, stats=None,
                 packed=None, init=None) -> SweepShard:
    """Sample datasets [lo, hi) on the current device and compute the K4 summaries (skip = python-side burn-in)."""
    from ._pack import pack
    from .engine import DeviceProblem
    if packed is None:
        ref_arr = np.broadcast_to(np.asarray(ref), (x.shape[0],))
        packed = pack(x[lo:hi], y[lo:hi], ref_arr[lo:hi], dtype=cfg.dtype)
    prob = DeviceProblem(packed, chains=cfg.chains)
    if init is None:
        init = initial_states(hi - lo, cfg.chains, packed.D, packed.K, cfg.seed, dataset0=lo)
    smp = prob.make_sampler(cfg.method, cfg.num_results, cfg.num_burnin, cfg.num_adapt, cfg.num_leapfrog,
                            cfg.max_tree_depth, cfg.step_size, cfg.target_accept, seed=cfg.seed,
                            chain_id0=lo * cfg.chains, warps_per_chain=cfg.warps_per_chain)
    out = prob.sample(init, smp, draws=draws, stats=stats)
    skip = min(cfg.num_burnin, cfg.num_results)          # second burn-in slice, scCODA_model.py:205-227
    summary = prob.summarize(out.draws, out.stats, skip)
    return SweepShard(lo=lo, hi=hi, problem=prob, out=out, summary=summary)


def shard_results(sh: SweepShard, cfg: SweepConfig) -> Dict[str, "object"]:
    """Fixed-size per-dataset results of a shard (device tensors, dim 0 = local datasets): what gets gathered."""
    import torch
    from . import diagnostics as dg
    prob = sh.problem
    s = prob.split_summary(sh.summary)
    n_used = float(cfg.num_results - min(cfg.num_burnin, cfg.num_results))
    B, C, D, K = prob.B, prob.C, prob.D, prob.K
    # pooled over the chains of a dataset (the reference has a single chain, result_classes.py:243)
    inc_prob = s["inc_count"].sum(1) / (C * n_used)
    cnt = s["inc_count"].sum(1)
    nz_mean = torch.where(cnt > 0, s["inc_sum"].sum(1) / cnt.clamp(min=1.0), torch.zeros_like(cnt))
    res = {
        "mean": s["mean"].mean(1),                                            # [B,P] posterior means
        "inclusion_prob": inc_prob,                                           # [B,D*K]
        "mean_nonzero": nz_mean,                                              # [B,D*K]
        "acc_rate": (s["n_accepted"].sum(1) / (C * float(cfg.num_results)))[:, None],
        "step_size": s["last_step_size"].mean(1)[:, None],
        "rhat_moments": dg.rhat_from_moments(s["mean"], s["m2"], int(n_used)) if C > 1 else
        torch.ones((B, prob.P), dtype=torch.float64, device=sh.summary.device),
    }
    nd = min(cfg.diag_datasets, B)
    if nd > 0 and C > 1:
        skip = min(cfg.num_burnin, cfg.num_results)
        ess = torch.full((B, K + D * (K - 1) + 1), float("nan"), dtype=torch.float64, device=sh.summary.device)
        rh = torch.full_like(ess, float("nan"))
        blk = 4
        for b0 in range(0, nd, blk):
            b1 = min(nd, b0 + blk)
            d = sh.out.draws[b0:b1, :, skip:, :].double()                      # [b,C,S,P]
            sig, boff = d[..., :D], d[..., D:D + D * (K - 1)]
            iraw, alpha = d[..., D + D * (K - 1):D + 2 * D * (K - 1)], d[..., prob.P - K:]
            beta = torch.sigmoid(50.0 * iraw) * boff * sig.repeat_interleave(K - 1, dim=-1)
            lp = sh.out.stats[b0:b1, :, skip:, 0:1].double()
            coords = torch.cat([alpha, beta, lp], dim=-1).permute(0, 3, 1, 2).contiguous()   # [b,coord,C,S]
            ess[b0:b1] = dg.ess_bulk(coords)
            rh[b0:b1] = dg.rhat(coords)
        res["ess_bulk"] = ess
        res["rhat_rank"] = rh
    if cfg.hdi_prob is not None:
        # HDIs as CAResult.summary_prepare reports them (result_classes.py:153-196), chains pooled: intercepts over all
        # retained draws, effects over b_raw * ind of the draws whose indicator is >= 1e-3 (NaN otherwise; the
        # reference column reads 0).  Device sort, a block of datasets at a time.
        skip = min(cfg.num_burnin, cfg.num_results)
        a_hdi = torch.zeros((B, K, 2), dtype=torch.float64, device=sh.summary.device)
        b_hdi = torch.zeros((B, D * K, 2), dtype=torch.float64, device=sh.summary.device)
        ref = torch.as_tensor(prob.packed.ref, device=sh.summary.device).long()
        nonref = torch.stack([torch.cat([torch.arange(0, int(r)), torch.arange(int(r) + 1, K)]) for r in prob.packed.ref]
                             ).to(sh.summary.device)                                        # [B,K-1] cell type of effect j
        blk = max(1, int(2 ** 27 // max(1, C * (cfg.num_results - skip) * (K + D * (K - 1)))))
        for b0 in range(0, B, blk):
            b1 = min(B, b0 + blk)
            d = sh.out.draws[b0:b1, :, skip:, :].to(torch.float32)                          # [b,C,S,P]
            nb_, S_ = d.shape[0], d.shape[2]
            alpha = d[..., prob.P - K:].permute(0, 3, 1, 2).reshape(nb_, K, C * S_)
            a_hdi[b0:b1] = dg.hdi(alpha, cfg.hdi_prob).double()
            sig = d[..., :D].repeat_interleave(K - 1, dim=-1)
            ind = torch.sigmoid(50.0 * d[..., D + D * (K - 1):D + 2 * D * (K - 1)])
            eff = torch.where(ind >= 1e-3, sig * d[..., D:D + D * (K - 1)] * ind, torch.full_like(ind, float("nan")))
            h = dg.hdi(eff.permute(0, 3, 1, 2).reshape(nb_, D * (K - 1), C * S_), cfg.hdi_prob).double()   # [b,D(K-1),2]
            for dd in range(D):
                idx = (dd * K + nonref[b0:b1])[:, :, None].expand(-1, -1, 2)
                b_hdi[b0:b1].scatter_(1, idx, h[:, dd * (K - 1):(dd + 1) * (K - 1)])
        res["alpha_hdi"] = a_hdi.reshape(B, K * 2)
        res["effect_hdi"] = b_hdi.reshape(B, D * K * 2)
    return res


def credible_effect_calls(inclusion_prob: np.ndarray, mean_nonzero: np.ndarray, est_fdr: float = 0.05):
    """Direct-posterior-probability threshold and final effects per dataset
    (sccoda/util/result_classes.py:262-283), vectorised over datasets on the host (tiny)."""
    B = inclusion_prob.shape[0]
    thresholds = np.ones(B)
    final = np.zeros_like(mean_nonzero)
    for b in range(B):
        inc = inclusion_prob[b]
        pos = np.sort(inc[inc > 0])[::-1]
        thr = 1.0
        for c in np.unique(pos):
            if np.mean(1.0 - pos[pos >= c]) < est_fdr:
                thr = np.floor(c * 1e3) / 1e3
                break
        thresholds[b] = thr
        final[b] = np.where(inc >= thr, mean_nonzero[b], 0.0)
    return thresholds, final


def run_sweep(x: np.ndarray, y: np.ndarray, ref, cfg: SweepConfig, comm: Optional[Comm] = None):
    """Shard B datasets over the ranks of `comm`, sample, summarise, gather.  Returns (results dict of numpy arrays
    on every rank, this rank's SweepShard)."""
    comm = comm or Comm()
    B = x.shape[0]
    if B < comm.world:
        # a rank without a dataset has nothing to pack or sample and no shapes to contribute to the gather
        raise ValueError(f"run_sweep needs at least one dataset per rank: {B} dataset(s) on {comm.world} ranks; "
                         f"launch at most {max(B, 1)} rank(s) for this sweep")
    keep = {}

    def compute(lo, hi):
        sh = sample_shard(x, y, ref, cfg, lo, hi)
        keep["shard"] = sh
        return shard_results(sh, cfg)

    gathered = run_sharded(B, compute, comm)
    res = {k: v.cpu().numpy() for k, v in gathered.items()}
    res["threshold"], res["final_effect"] = credible_effect_calls(res["inclusion_prob"], res["mean_nonzero"],
                                                                  cfg.est_fdr)
    return res, keep["shard"]
