"""
Host-side packer: raw (x, y, reference) -> the packed dataset layout of include/sccoda_b200.h.

`pack` is a thin binding of the library's own packer (`sccoda_pack`, csrc/pack.cpp: C++, threaded over datasets —
the one `sccoda_sample_raw_host` runs inside the end-to-end call).  `pack_numpy` is the same layout written in
numpy; it documents the layout, and tests/test_pack_native.py holds the two to each other.

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
    # ---- grouped ("pair / item") layout of the gradient pass, see build_items() --------------------------------
    grouped: int = 0                 # 1: the kernels take the pair/item gradient path
    NImax: int = 0                   # item slots per dataset (even)
    NOmax: int = 0                   # odd-count slots per dataset (>= 1)
    NGmax: int = 0                   # odd-group slots per dataset (>= 1)
    NF: int = 0                      # result slots per pair (power of two >= 2): its items + its odd group
    NT: int = 0                      # largest number of items of one pair
    NOT: int = 0                     # largest number of odd counts of one pair
    NI: np.ndarray = None            # [B] int32 items of dataset b
    NG: np.ndarray = None            # [B] int32 odd groups of dataset b
    iy: np.ndarray = None            # [B,ITEM_R,NImax] counts >= 6 of each item, padded with 6
    ipair: np.ndarray = None         # [B,NImax] uint16 pair of each item
    ipos: np.ndarray = None          # [B,NImax] uint16 result slot of each item: pair*NF + t
    inreal: np.ndarray = None        # [B,NImax] uint8 counts of the item that are data (the rest is padding)
    pcoef: np.ndarray = None         # [B,NP,6] rational coefficients of each pair, NP = Umax*(K+1)
    oy: np.ndarray = None            # [B,NOmax] counts < 6 that are not in {0,1,...,5}
    opair: np.ndarray = None         # [B,NGmax] uint16 pair of each odd group
    opos: np.ndarray = None          # [B,NGmax] uint16 result slot of each odd group
    odesc: np.ndarray = None         # [B,NGmax] uint32 first odd count | n_odd << 16

    @property
    def NP(self) -> int:
        return self.Umax * (self.K + 1)

    @property
    def P(self) -> int:
        return self.D + 2 * self.D * (self.K - 1) + self.K


def apply_pseudocount(y: np.ndarray) -> np.ndarray:
    """scCODA_model.py:94-96 — only when at least one zero is present; 0 -> 0.5."""
    y = np.array(y, dtype=np.float64)
    if np.count_nonzero(y) != y.size:
        y[y == 0] = 0.5
    return y


ITEM_R = 5  # counts per item (SCCODA_ITEM_R in include/sccoda_b200.h)


def build_items(ys: np.ndarray, ntot: np.ndarray, rstart: np.ndarray, U: np.ndarray, Umax: int):
    """Pair / item layout of the gradient pass (csrc/grouped.cuh).

    A *pair* is (covariate group u, cell type k): all rows of the group share the concentration c_uk, so
    sum_{n in u} [psi(c + y_nk) - psi(c)] only depends on the multiset of counts of the pair.  The row totals form
    one more pair per group ("total pair", index Umax*K + u, concentration S_u = sum_k c_uk).  Per pair:
      * counts >= 6 go, ITEM_R at a time, into *items* (padded with 6, which contributes psi(c+6) - psi(c+6) = 0);
        an item yields sum_j [psi(c + y_j) - psi(c + 6)] with one log of a product and no recurrence shift;
      * counts in {1,...,5} and the items' reference point psi(c+6) - psi(c) = sum_{i<6} 1/(c+i) are folded into
        six weights W_i >= 0 of sum_i W_i / (c + i), stored as the numerators of three quadratics
        (c(c+5), (c+1)(c+4), (c+2)(c+3));  counts equal to 0 contribute nothing;
      * any other count below 6 (the 0.5 pseudocount, non-integer data) is an *odd* count, evaluated one by one;
        the odd counts of one pair form an *odd group*.
    Every item / odd group owns one of the NF result slots of its pair (slot index pair*NF + t).
    Returns a dict of float64 / integer arrays (see include/sccoda_b200.h for the layout).
    """
    B, N, K = ys.shape
    R = ITEM_R
    NP = Umax * (K + 1)
    ext = np.concatenate([ys, ntot[:, :, None]], axis=2)                      # [B,N,K+1]: last column = row totals
    pcoef = np.zeros((B, NP, 6))
    per_ds = []
    nslot_max = 1
    nt_max, not_max = 1, 0
    for b in range(B):
        yb, pb, tb, rb, ob, opb, otb, odb = [], [], [], [], [], [], [], []
        n_odd = 0
        for u in range(int(U[b])):
            blk = ext[b, rstart[b, u]:rstart[b, u + 1]]                       # [cnt,K+1]
            pair = np.where(np.arange(K + 1) < K, u * K + np.arange(K + 1), Umax * K + u)
            big = blk >= 6.0
            is_int = (blk == np.floor(blk)) & (blk >= 1.0) & ~big
            odd = ~big & ~is_int & (blk != 0.0)
            nb = big.sum(axis=0)
            no = odd.sum(axis=0)
            W = np.stack([nb + no + (is_int & (blk > i)).sum(axis=0) for i in range(6)], axis=1).astype(np.float64)
            pcoef[b, pair] = np.stack([W[:, 0] + W[:, 5], 5.0 * W[:, 0], W[:, 1] + W[:, 4], 4.0 * W[:, 1] + W[:, 4],
                                       W[:, 2] + W[:, 3], 3.0 * W[:, 2] + 2.0 * W[:, 3]], axis=1)
            nf = -(-nb // R)
            nslot_max = max(nslot_max, int((nf + (no > 0)).max()))
            nt_max, not_max = max(nt_max, int(nf.max())), max(not_max, int(no.max()))
            # columns sorted by decreasing count: the big ones come first, everything else becomes padding
            srt = -np.sort(-np.where(big, blk, 6.0), axis=0)
            depth = int(nf.max()) * R
            if depth > srt.shape[0]:
                srt = np.concatenate([srt, np.full((depth - srt.shape[0], K + 1), 6.0)], axis=0)
            if depth:
                chunks = srt[:depth].reshape(depth // R, R, K + 1)            # [t, j, col]
                tt, cc = np.nonzero(np.arange(depth // R)[:, None] < nf[None, :])
                order = np.lexsort((tt, cc))                                  # pair-major, items of a pair consecutive
                tt, cc = tt[order], cc[order]
                yb.append(chunks[tt, :, cc])
                pb.append(pair[cc])
                tb.append(tt)
                rb.append(np.minimum(R, nb[cc] - R * tt))                     # real (non-padding) counts of the item
            if no.any():
                rr, cc = np.nonzero(odd.T)                                    # rr = column, cc = row (column-major)
                ob.append(blk[cc, rr])
                cols = np.nonzero(no)[0]
                opb.append(pair[cols])
                otb.append(nf[cols])                                          # slot of the group: after the items
                odb.append((n_odd + np.r_[0, np.cumsum(no[cols])[:-1]]) | (no[cols] << 16))
                n_odd += int(no.sum())
        cat = lambda lst, shape, dt: np.concatenate(lst).astype(dt) if lst else np.zeros(shape, dt)
        per_ds.append((cat(yb, (0, R), np.float64), cat(pb, (0,), np.int64), cat(tb, (0,), np.int64),
                       cat(ob, (0,), np.float64), cat(opb, (0,), np.int64), cat(otb, (0,), np.int64),
                       cat(odb, (0,), np.int64), cat(rb, (0,), np.int64)))
    NF = 2
    while NF < nslot_max:
        NF *= 2
    NI = np.array([d[1].size for d in per_ds], dtype=np.int32)
    NG = np.array([d[4].size for d in per_ds], dtype=np.int32)
    NImax = int(NI.max()) if B else 0
    NImax = max(2, NImax + (NImax & 1))
    NOmax = max(1, max(d[3].size for d in per_ds) if B else 1)
    NGmax = max(1, int(NG.max()) if B else 1)
    if NP * NF > 65535 or NImax > 65535 or NOmax > 65535:
        raise ValueError("pair/item layout needs Umax*(K+1)*NF <= 65535 and at most 65535 items / odd counts")
    iy = np.full((B, R, NImax), 6.0)
    ipair = np.zeros((B, NImax), dtype=np.uint16)
    ipos = np.full((B, NImax), NP * NF, dtype=np.uint16)                      # padding items write to a spare slot
    oy = np.full((B, NOmax), 1.0)
    opair = np.zeros((B, NGmax), dtype=np.uint16)
    opos = np.full((B, NGmax), NP * NF, dtype=np.uint16)
    odesc = np.zeros((B, NGmax), dtype=np.uint32)
    inreal = np.zeros((B, NImax), dtype=np.uint8)
    for b, (yb, pb, tb, ob, opb, otb, odb, rb) in enumerate(per_ds):
        iy[b, :, :NI[b]] = yb.T
        ipair[b, :NI[b]] = pb
        inreal[b, :NI[b]] = rb
        ipos[b, :NI[b]] = pb * NF + tb
        oy[b, :ob.size] = ob
        opair[b, :NG[b]] = opb
        opos[b, :NG[b]] = opb * NF + otb
        odesc[b, :NG[b]] = odb
    return dict(NT=nt_max, NOT=not_max, NF=NF, NImax=NImax, NOmax=NOmax, NGmax=NGmax, NI=NI, NG=NG, iy=iy, ipair=ipair, ipos=ipos, pcoef=pcoef,
                oy=oy, opair=opair, opos=opos, odesc=odesc, inreal=inreal)


def _no_items(B: int, K: int, Umax: int) -> dict:
    """Placeholder pair/item arrays for a problem that only has the element layout."""
    NP = Umax * (K + 1)
    return dict(NT=1, NOT=0, NF=2, NImax=2, NOmax=1, NGmax=1, NI=np.zeros(B, np.int32), NG=np.zeros(B, np.int32),
                iy=np.full((B, ITEM_R, 2), 6.0), ipair=np.zeros((B, 2), np.uint16), ipos=np.zeros((B, 2), np.uint16),
                pcoef=np.zeros((B, NP, 6)), oy=np.ones((B, 1)), opair=np.zeros((B, 1), np.uint16),
                opos=np.zeros((B, 1), np.uint16), odesc=np.zeros((B, 1), np.uint32), inreal=np.zeros((B, 2), np.uint8))


def _normalise_inputs(x, y, ref):
    """Shapes and errors shared by both packers (the reference's own checks: scCODA_model.py:110-113, comp_ana.py:129-130)."""
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
    ref = np.broadcast_to(np.asarray(ref, dtype=np.int32), (B,)).copy()
    if np.any(ref < 0) or np.any(ref >= K):
        raise NameError("Reference index is not a valid cell type name or numerical index!")
    return np.ascontiguousarray(x), np.ascontiguousarray(y), ref


def pack(x, y, ref, dtype=np.float32, pseudocount: bool = True, grouped=None, threads: int = 0) -> PackedProblem:
    """x: [B,N,D] or [N,D]; y: [B,N,K] or [N,K]; ref: int or [B].  Runs the library's packer (`sccoda_pack`) and copies
    its arrays into a PackedProblem; arguments as in `pack_numpy`.  Raises when the native library is missing."""
    import ctypes
    from . import _native as nat
    x, y, ref = _normalise_inputs(x, y, ref)
    B, N, K = y.shape
    D = x.shape[2]
    dtype = np.dtype(dtype)
    lib = nat.lib()
    handle = ctypes.c_void_p()
    rc = lib.sccoda_pack(B, N, K, D, x.ctypes.data, y.ctypes.data, ref.ctypes.data, nat.F32 if dtype.itemsize == 4 else nat.F64,
                         int(bool(pseudocount)), -1 if grouped is None else int(bool(grouped)), int(threads),
                         ctypes.byref(handle))
    if rc in (-53, -55):
        raise ValueError(lib.sccoda_last_error().decode("utf-8", "replace"))
    nat.check(rc, "sccoda_pack")
    try:
        prob = nat.Problem()
        nat.check(lib.sccoda_packed_problem(handle, 1, ctypes.byref(prob)), "sccoda_packed_problem")
        Umax, NP = prob.Umax, prob.Umax * (K + 1)

        def arr(addr, shape, dt):
            n = int(np.prod(shape))
            buf = (ctypes.c_char * (n * np.dtype(dt).itemsize)).from_address(addr)
            return np.frombuffer(buf, dtype=dt, count=n).reshape(shape).copy()

        ys = arr(lib.sccoda_packed_y(handle), (B, N, K), np.float64)
        pk = PackedProblem(
            B=B, N=N, K=K, D=D, Umax=Umax, dtype=dtype,
            ey=arr(prob.ey, (B, N * K), dtype), eycst=arr(prob.eycst, (B, N * K), dtype),
            eidx=arr(prob.eidx, (B, N * K), np.uint32), colperm=arr(prob.colperm, (B, K), np.int32),
            nbig=arr(prob.nbig, (B,), np.int32), nbig_min=prob.nbig_min, y=ys.astype(dtype),
            ntot=arr(prob.ntot, (B, N), dtype), xu=arr(prob.xu, (B, Umax, D), dtype), grp=arr(prob.grp, (B, N), np.int32),
            rstart=arr(prob.rstart, (B, Umax + 1), np.int32), U=arr(prob.U, (B,), np.int32), ref=arr(prob.ref, (B,), np.int32),
            lconst=arr(prob.lconst, (B,), np.float64),
            row_order=arr(lib.sccoda_packed_row_order(handle), (B, N), np.int32).astype(np.int64),
            logp_const=arr(lib.sccoda_packed_logp_const(handle), (B,), np.float64),
            grouped=prob.grouped, NImax=prob.NImax, NOmax=prob.NOmax, NGmax=prob.NGmax, NF=prob.NF, NT=prob.NT, NOT=prob.NOT,
            NI=arr(prob.NI, (B,), np.int32), NG=arr(prob.NG, (B,), np.int32), iy=arr(prob.iy, (B, ITEM_R, prob.NImax), dtype),
            ipair=arr(prob.ipair, (B, prob.NImax), np.uint16), ipos=arr(prob.ipos, (B, prob.NImax), np.uint16),
            inreal=arr(prob.inreal, (B, prob.NImax), np.uint8), pcoef=arr(prob.pcoef, (B, NP, 6), dtype),
            oy=arr(prob.oy, (B, prob.NOmax), dtype), opair=arr(prob.opair, (B, prob.NGmax), np.uint16),
            opos=arr(prob.opos, (B, prob.NGmax), np.uint16), odesc=arr(prob.odesc, (B, prob.NGmax), np.uint32))
    finally:
        lib.sccoda_packed_free(handle)
    return pk


def pack_numpy(x, y, ref, dtype=np.float32, pseudocount: bool = True, grouped=None) -> PackedProblem:
    """x: [B,N,D] or [N,D]; y: [B,N,K] or [N,K]; ref: int or [B].

    grouped: None = choose (pair/item gradient path for the designs the lane-per-cell-type kernels take — at most 32
    covariate groups, D <= 4 and K <= 31 or D <= 3 and K <= 62 —
    and whenever the covariate groups are large enough that the item padding stays below 60 % of the counts — unless
    the pair/item arrays of the dataset would not fit one CTA's shared memory), True / False = force."""
    args_in, auto = (x, y, ref), grouped is None
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
        with np.errstate(divide="ignore", invalid="ignore"):    # exact zeros (pseudocount=False) are never "big"
            stirling = (ys - 0.5) * np.log(ys) - ys + 0.5 * LOG_2PI
            ycst = np.where(big, lg - stirling, lg)             # r(y) for y >= 6, lgamma(y) otherwise
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
    try:
        it = build_items(ys, ntot, rstart, U, Umax)
    except ValueError:
        if grouped:
            raise
        # too many pairs / items for the 16-bit indices of the pair/item layout: element layout, empty item arrays
        it = _no_items(B, K, Umax)
        grouped = False
    if grouped is None:
        # designs with few covariate groups (case/control, a handful of conditions, small factorials) always take the
        # pair/item layout: the lane-per-cell-type kernels (hmc_cc.cuh, hmc_mg.cuh) evaluate a padded item as cheaply
        # as a full one; otherwise only when the item padding stays below 60 % of the counts
        lane_owned = Umax <= 32 and ((K <= 31 and D <= 4) or (K <= 62 and D <= 3))
        grouped = bool(lane_owned or ITEM_R * int(it["NI"].max()) <= 1.6 * N * (K + 1))
    if grouped:
        # The pair/item gradient does not walk the element list; only the value pass does, and it does not care
        # about the order.  List the counts >= 6 first (stable), so that the fast loop covers all of them and not
        # just the columns without a small count.
        order = np.argsort(ey < 6.0, axis=1, kind="stable")
        ey, eycst, eidx = (np.take_along_axis(a, order, axis=1) for a in (ey, eycst, eidx))
        nbig = (ey >= 6.0).sum(axis=1).astype(np.int32)
    pk = PackedProblem(grouped=int(bool(grouped)), NImax=it["NImax"], NOmax=it["NOmax"], NGmax=it["NGmax"],
                         NF=it["NF"], NT=it["NT"], NOT=it["NOT"], NI=it["NI"], NG=it["NG"], iy=np.ascontiguousarray(it["iy"], dtype),
                         ipair=it["ipair"], ipos=it["ipos"], inreal=it["inreal"], pcoef=np.ascontiguousarray(it["pcoef"], dtype),
                         oy=np.ascontiguousarray(it["oy"], dtype), opair=it["opair"], opos=it["opos"],
                         odesc=it["odesc"],
                         B=B, N=N, K=K, D=D, Umax=Umax, dtype=dtype,
                         ey=np.ascontiguousarray(ey, dtype), eycst=np.ascontiguousarray(eycst, dtype),
                         eidx=np.ascontiguousarray(eidx), colperm=np.ascontiguousarray(colperm),
                         nbig=nbig, nbig_min=int(nbig.min()),
                         y=np.ascontiguousarray(ys, dtype),
                         ntot=np.ascontiguousarray(ntot, dtype), xu=np.ascontiguousarray(xu, dtype),
                         grp=np.ascontiguousarray(grp), rstart=np.ascontiguousarray(rstart), U=U, ref=ref,
                         lconst=np.ascontiguousarray(lconst, np.float64),
                         row_order=np.stack(orders), logp_const=logp_const)
    if auto and pk.grouped and not _grouped_layout_fits(pk):
        return pack_numpy(*args_in, dtype=dtype, pseudocount=pseudocount, grouped=False)
    return pk


def _grouped_layout_fits(pk: PackedProblem) -> bool:
    """Can the kernels hold the pair/item layout of this problem in one CTA's shared memory?  Asked of the library's
    own launch planner (`sccoda_launch_plan`, NUTS with the default tree depth = the largest footprint); very long
    datasets (thousands of rows) fall back to the element layout, whose list streams from global memory."""
    import ctypes
    try:
        from . import _native as nat
        lib = nat.lib()
    except (OSError, RuntimeError):
        return True                                  # no library here: sampling will complain loudly, not the packer
    prob = nat.Problem(C=1, **nat.problem_fields(pk, lambda name: getattr(pk, name).ctypes.data))
    smp = nat.Sampler(method=nat.METHOD_NUTS, num_results=1, num_burnin=0, num_adapt=0, num_leapfrog=10, max_tree_depth=10,
                      step_begin=0, step_end=0, step_size=0.01, target_accept=0.75, seed=0, chain_id0=0,
                      warps_per_chain=0, store_draws=1, chains_per_cta=0, generic_only=0, freeze_spike_slab=0,
                      reject_conc_above=0.0)
    out = [ctypes.c_int32() for _ in range(4)]
    rc = lib.sccoda_launch_plan(ctypes.byref(prob), ctypes.byref(smp), *(ctypes.byref(v) for v in out))
    return rc != -9
