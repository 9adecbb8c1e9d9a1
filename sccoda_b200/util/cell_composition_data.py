"""
Minimal compositional-data container and importers.

The reference passes `anndata.AnnData` objects around (sccoda/util/cell_composition_data.py:15-296,
sccoda/util/comp_ana.py:67-75).  anndata is an optional dependency here: `CompositionData` offers the three
attributes the sampler path reads (`X` counts [samples x cell types], `obs` covariate DataFrame, `var` with the
cell-type index) plus `uns`; real AnnData objects are accepted everywhere a CompositionData is.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import pandas as pd


class CompositionData:
    """Duck-typed stand-in for AnnData: X [N,K], obs (DataFrame, N rows), var (DataFrame indexed by cell type)."""

    def __init__(self, X, obs: Optional[pd.DataFrame] = None, var: Optional[pd.DataFrame] = None,
                 uns: Optional[dict] = None):
        self.X = np.asarray(X)
        n, k = self.X.shape
        if obs is None:
            obs = pd.DataFrame(index=pd.RangeIndex(n).astype(str))
        if var is None:
            var = pd.DataFrame(index=pd.RangeIndex(k).astype(str))
        if len(obs) != n or len(var) != k:
            raise ValueError(f"X is {self.X.shape} but obs has {len(obs)} rows and var has {len(var)} rows")
        self.obs = obs
        self.var = var
        self.uns = {} if uns is None else uns

    @property
    def shape(self):
        return self.X.shape

    @property
    def n_obs(self):
        return self.X.shape[0]

    @property
    def n_vars(self):
        return self.X.shape[1]

    def copy(self):
        return CompositionData(self.X.copy(), self.obs.copy(), self.var.copy(), dict(self.uns))

    def __getitem__(self, idx):
        """Row subsetting with a boolean mask / index list / obs-label list (what the tutorials use)."""
        if isinstance(idx, tuple):
            rows, cols = idx
        else:
            rows, cols = idx, slice(None)
        if isinstance(rows, pd.Series):
            rows = rows.to_numpy()
        rows = np.asarray(rows) if not isinstance(rows, slice) else rows
        if not isinstance(rows, slice) and rows.dtype.kind in "OUS":
            rows = self.obs.index.get_indexer(rows)
        ridx = np.arange(self.n_obs)[rows]
        if isinstance(cols, slice):
            cidx = np.arange(self.n_vars)[cols]
        else:
            cols = np.asarray(cols)
            cidx = self.var.index.get_indexer(cols) if cols.dtype.kind in "OUS" else np.arange(self.n_vars)[cols]
        return CompositionData(self.X[np.ix_(ridx, cidx)], self.obs.iloc[ridx].copy(), self.var.iloc[cidx].copy(),
                               dict(self.uns))

    def __repr__(self):
        return (f"CompositionData object with n_obs x n_vars = {self.n_obs} x {self.n_vars}\n"
                f"    obs: {list(self.obs.columns)}\n    var: {list(self.var.columns)}")


def from_pandas(df: pd.DataFrame, covariate_columns: List[str]) -> CompositionData:
    """Counts + covariates in one DataFrame -> CompositionData (cf. sccoda/util/cell_composition_data.py:263-296).

    Every column not listed in `covariate_columns` is a cell type; row labels become strings.
    """
    covariate_data = df.loc[:, covariate_columns].copy()
    covariate_data.index = covariate_data.index.astype(str)
    count_cols = [c for c in df.columns if c not in set(covariate_columns)]
    counts = df.loc[:, count_cols]
    var = pd.DataFrame(index=pd.Index([str(c) for c in count_cols]))
    return CompositionData(counts.to_numpy(dtype=np.float64), covariate_data, var)
