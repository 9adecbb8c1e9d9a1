"""
Host-side packer: raw (x, y, reference) -> the packed dataset layout of include/sccoda_b200.h.

Mirrors the input handling of CompositionalModel.__init__ (sccoda/model/scCODA_model.py:90-100): counts are
cast to float64, zeros are replaced by a 0.5 pseudocount BEFORE the row sums are taken.  Rows are then sorted
by covariate group (rows with identical covariate vectors share their concentrations, see csrc/model.cuh), the
counts are laid out as a column-major "element list" with the cell types sorted by decreasing minimum count
(so that the elements whose digamma needs no recurrence shift come first), and every parameter-independent
term of the log-density is folded into one float64 constant per dataset.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
from scipy.special import gammaln

LOG_2PI = math.log(2.0 * math.pi)


@dataclass
class PackedProblem:
    """B datasets of equal shape, host arrays in the C-ABI layout (dtype float32 or float64)."""
    B: int
    N: int
    K: int
    D: int
    Umax: int
    dtype: np.dtype
    ey: np.ndarray       # [B,N*K]   counts, element-list order (e = kk*N + n)
    eycst: np.ndarray    # [B,N*K]
    eidx: np.ndarray     # [B,N*K] uint32  (u*K + k) | (n << 16)
    colperm: np.ndarray  # [B,K] int32  sorted column position -> original cell type
    nbig: np.ndarray     # [B] int32  leading elements with counts >= 6
    nbig_min: int
    y: np.ndarray        # [B,N,K]   counts after pseudocount, rows sorted by group (host-side reference copy)
    ntot: np.ndarray     # [B,N]
    xu: np.ndarray       # [B,Umax,D]
    grp: np.ndarray      # [B,N] int32
    rstart: np.ndarray   # [B,Umax+1] int32
    U: np.ndarray        # [B] int32
    ref: np.ndarray      # [B] int32
    lconst: np.ndarray   # [B] float64
    row_order: np.ndarray  # [B,N] original row index of each packed row
    logp_const: np.ndarray  # [B] multinomial-coefficient constant (float64), for reference

    @property
    def P(self) -> int:
        return self.D + 2 * self.D * (self.K - 1) + self.K


def apply_pseudocount(y: np.ndarray) -> np.ndarray:
    """scCODA_model.py:94-96 — only when at least one zero is present; 0 -> 0.5."""
    y = np.array(y, dtype=np.float64)
    if np.count_nonzero(y) != y.size:
        y[y == 0] = 0.5
    return y


def pack(x, y, ref, dtype=np.float32, pseudocount: bool = True) -> PackedProblem:
    """x: [B,N,D] or [N,D]; y: [B,N,K] or [N,K]; ref: int or [B]."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    if y.ndim == 2:
        y = y[None]
    if x.ndim == 1:
        x = x[:, None]
    if x.ndim == 2:
        x = np.broadcast_to(x[None], (y.shape[0],) + x.shape)
    B, N, K = y.shape
    if x.shape[0] != B or x.shape[1] != N:
        raise ValueError("Wrong input dimensions X[{},:] != y[{},:]".format(x.shape[1], N))
    D = x.shape[2]
    ref = np.broadcast_to(np.asarray(ref, dtype=np.int32), (B,)).copy()
    if np.any(ref < 0) or np.any(ref >= K):
        raise NameError("Reference index is not a valid cell type name or numerical index!")
    dtype = np.dtype(dtype)
    if pseudocount:
        y = np.stack([apply_pseudocount(y[b]) for b in range(B)])

    orders, groups, uniq = [], [], []
    for b in range(B):
        u_rows, inv = np.unique(x[b], axis=0, return_inverse=True)
        inv = np.asarray(inv).reshape(-1)
        order = np.argsort(inv, kind="stable")
        orders.append(order)
        groups.append(inv[order].astype(np.int32))
        uniq.append(u_rows)
    Umax = max(u.shape[0] for u in uniq)

    ys = np.stack([y[b][orders[b]] for b in range(B)])
    ntot = ys.sum(axis=2)                                        # :99-100 (after the pseudocount)
    xu = np.zeros((B, Umax, D))
    rstart = np.full((B, Umax + 1), N, dtype=np.int32)
    U = np.zeros(B, dtype=np.int32)
    for b in range(B):
        u = uniq[b].shape[0]
        U[b] = u
        xu[b, :u] = uniq[b]
        rstart[b, :u] = np.searchsorted(groups[b], np.arange(u), side="left")
    grp = np.stack(groups).astype(np.int32)

    logp_const = gammaln(ntot + 1.0).sum(axis=1) - gammaln(ys + 1.0).sum(axis=(1, 2))
    DK1 = D * (K - 1)
    prior_const = D * math.log(2.0 / math.pi) - DK1 * LOG_2PI - K * (math.log(5.0) + 0.5 * LOG_2PI)
    if dtype == np.float32:
        lg = gammaln(ys)
        big = ys >= 6.0
        stirling = (ys - 0.5) * np.log(ys) - ys + 0.5 * LOG_2PI
        ycst = np.where(big, lg - stirling, lg)                 # r(y) for y >= 6, lgamma(y) otherwise
        lconst = prior_const + logp_const + lg.sum(axis=(1, 2))
    else:
        ycst = np.zeros_like(ys)
        lconst = prior_const + logp_const
    if Umax * K > 65535 or N > 65535:
        raise ValueError("packed element index needs Umax*K <= 65535 and N <= 65535")
    # element list: column-major, columns sorted by decreasing minimum count (stable)
    colmin = ys.min(axis=1)                                                   # [B,K]
    colperm = np.argsort(-colmin, axis=1, kind="stable").astype(np.int32)     # [B,K]
    nbig = (N * (colmin >= 6.0).sum(axis=1)).astype(np.int32)
    bidx = np.arange(B)[:, None, None]
    perm = colperm[:, :, None]                                                # [B,K,1]
    rows = np.arange(N)[None, None, :]                                        # [1,1,N]
    ey = ys[bidx, rows, perm].reshape(B, N * K)                               # [B,K,N] -> e = kk*N + n
    eycst = ycst[bidx, rows, perm].reshape(B, N * K)
    uk = grp[:, None, :].astype(np.int64) * K + perm                          # [B,K,N]
    eidx = (uk | (rows.astype(np.int64) << 16)).astype(np.uint32).reshape(B, N * K)
    return PackedProblem(B=B, N=N, K=K, D=D, Umax=Umax, dtype=dtype,
                         ey=np.ascontiguousarray(ey, dtype), eycst=np.ascontiguousarray(eycst, dtype),
                         eidx=np.ascontiguousarray(eidx), colperm=np.ascontiguousarray(colperm),
                         nbig=nbig, nbig_min=int(nbig.min()),
                         y=np.ascontiguousarray(ys, dtype),
                         ntot=np.ascontiguousarray(ntot, dtype), xu=np.ascontiguousarray(xu, dtype),
                         grp=np.ascontiguousarray(grp), rstart=np.ascontiguousarray(rstart), U=U, ref=ref,
                         lconst=np.ascontiguousarray(lconst, np.float64),
                         row_order=np.stack(orders), logp_const=logp_const)
