"""
Cross-chain convergence diagnostics on device: rank-normalised split-R-hat and bulk ESS (Vehtari et al. 2021,
the arviz / Stan definitions named in SURVEY §8d).  The reference computes neither (it runs one chain and calls
az.summary(kind="stats"), sccoda/util/result_classes.py:153); these are the new multi-chain capability.

All functions take a torch tensor x[..., C, S] (C chains, S draws; any leading batch dims, e.g. datasets x
coordinates) on any device and reduce the last two dims.  torch is used for sort / FFT (library post-processing,
not the sampler hot path).
"""
from __future__ import annotations

import math

import torch


def _split_chains(x: torch.Tensor) -> torch.Tensor:
    """[..., C, S] -> [..., 2C, S//2] (drops the middle draw when S is odd)."""
    S = x.shape[-1]
    h = S // 2
    return torch.cat([x[..., :h], x[..., S - h:]], dim=-2)


def _z_scale(x: torch.Tensor) -> torch.Tensor:
    """Rank-normalise over the pooled (chain, draw) sample: z = Phi^-1((rank - 3/8) / (n + 1/4)), average ranks
    are approximated by ordinal ranks (ties have measure zero for continuous draws)."""
    shp = x.shape
    flat = x.reshape(*shp[:-2], -1)
    n = flat.shape[-1]
    order = flat.argsort(dim=-1)
    ranks = torch.empty_like(order)
    src = torch.arange(1, n + 1, device=x.device).expand_as(order)
    ranks.scatter_(-1, order, src)
    p = (ranks.to(torch.float64) - 0.375) / (n + 0.25)
    z = math.sqrt(2.0) * torch.erfinv(2.0 * p - 1.0)
    return z.reshape(shp)


def _autocov(x: torch.Tensor) -> torch.Tensor:
    """Biased autocovariance along the last dim via FFT (same normalisation as arviz.stats.autocov)."""
    S = x.shape[-1]
    n_fft = 1 << (2 * S - 1).bit_length()
    xc = x - x.mean(dim=-1, keepdim=True)
    f = torch.fft.rfft(xc, n=n_fft, dim=-1)
    ac = torch.fft.irfft(f * f.conj(), n=n_fft, dim=-1)[..., :S]
    return ac / S


def _ess_from_z(z: torch.Tensor) -> torch.Tensor:
    """ESS of already transformed draws z[..., C, S] (Geyer initial positive + monotone sequence)."""
    C, S = z.shape[-2], z.shape[-1]
    acov = _autocov(z)                                            # [..., C, S]
    chain_mean = z.mean(dim=-1)
    mean_var = acov[..., 0].mean(dim=-1) * S / (S - 1.0)
    var_plus = mean_var * (S - 1.0) / S
    if C > 1:
        var_plus = var_plus + chain_mean.var(dim=-1, unbiased=True)
    rho = 1.0 - (mean_var[..., None] - acov.mean(dim=-2)) / var_plus[..., None]   # [..., S]
    rho[..., 0] = 1.0
    n_pairs = S // 2
    pairs = rho[..., : 2 * n_pairs].reshape(*rho.shape[:-1], n_pairs, 2).sum(dim=-1)   # P_j = rho_2j + rho_2j+1
    positive = pairs > 0
    positive[..., 0] = True
    alive = torch.cumprod(positive.to(torch.int8), dim=-1).bool()                       # until the first P_j <= 0
    mono = torch.cummin(pairs, dim=-1).values
    tau = -1.0 + 2.0 * torch.where(alive, mono, torch.zeros_like(mono)).sum(dim=-1)
    # "improve estimation": add the first even-lag autocorrelation after the cut when it is positive
    n_alive = alive.sum(dim=-1)                                                         # >= 1
    idx = (2 * n_alive).clamp(max=S - 1)
    extra = rho.gather(-1, idx[..., None])[..., 0]
    tau = tau + torch.where((extra > 0) & (n_alive < n_pairs), extra, torch.zeros_like(extra))
    ess_max = float(C * S)
    tau = torch.clamp(tau, min=1.0 / math.log10(ess_max))
    return ess_max / tau


def ess_bulk(x: torch.Tensor) -> torch.Tensor:
    """Bulk effective sample size of x[..., C, S] -> [...]."""
    return _ess_from_z(_z_scale(_split_chains(x.to(torch.float64))))


def _rhat_plain(z: torch.Tensor) -> torch.Tensor:
    S = z.shape[-1]
    chain_mean = z.mean(dim=-1)
    chain_var = z.var(dim=-1, unbiased=True)
    between = S * chain_mean.var(dim=-1, unbiased=True)      # B = S * var(chain means)
    within = chain_var.mean(dim=-1)
    return torch.sqrt((between / within + S - 1.0) / S)


def rhat(x: torch.Tensor) -> torch.Tensor:
    """Rank-normalised split-R-hat (max of the bulk and the folded/tail version) of x[..., C, S] -> [...]."""
    xs = _split_chains(x.to(torch.float64))
    bulk = _rhat_plain(_z_scale(xs))
    med = xs.reshape(*xs.shape[:-2], -1).median(dim=-1).values
    folded = (xs - med[..., None, None]).abs()
    tail = _rhat_plain(_z_scale(folded))
    return torch.maximum(bulk, tail)


def rhat_from_moments(mean: torch.Tensor, m2: torch.Tensor, n: int) -> torch.Tensor:
    """Classic (non-split) R-hat from per-chain moments: mean[..., C, P], m2[..., C, P] (sum of squared deviations),
    n draws per chain -> [..., P].  This is what can be computed from the tiny summaries gathered over NCCL."""
    within = (m2 / (n - 1.0)).mean(dim=-2)
    between = mean.var(dim=-2, unbiased=True)
    return torch.sqrt(((n - 1.0) / n * within + between) / within)


def hdi(x: torch.Tensor, prob: float = 0.94) -> torch.Tensor:
    """Highest-density interval of the draws along the last dim, for every leading index at once: the shortest
    interval that contains floor(prob * n) of the sorted non-NaN draws — what az.summary(kind="stats", skipna=True)
    reports for a unimodal sample and what CAResult.summary_prepare reads (sccoda/util/result_classes.py:153-196;
    NaN marks draws whose indicator is below 1e-3, :178).  Returns [..., 2] (NaN where no draw is valid)."""
    v, _ = torch.sort(x, dim=-1)                                   # NaNs sort to the end
    n = (~torch.isnan(x)).sum(dim=-1, keepdim=True)                # valid draws per row
    k = torch.floor(prob * n.to(torch.float64)).to(torch.long)     # window length (in steps) per row
    S = x.shape[-1]
    pos = torch.arange(S, device=x.device).expand(x.shape)
    hi_idx = torch.clamp(pos + k, max=S - 1)
    width = torch.gather(v, -1, hi_idx) - v
    width = torch.where(pos + k < n, width, torch.full_like(width, float("inf")))
    best = torch.argmin(width, dim=-1, keepdim=True)
    lo = torch.gather(v, -1, best)
    hi = torch.gather(v, -1, torch.clamp(best + k, max=S - 1))
    out = torch.cat([lo, hi], dim=-1)
    empty = (n == 0) | (k >= n)
    return torch.where(empty.expand_as(out), torch.full_like(out, float("nan")), out)
