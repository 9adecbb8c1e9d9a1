"""
Result object of a compositional analysis — API-compatible stand-in for `sccoda.util.result_classes.CAResult`
(sccoda/util/result_classes.py:47-619).

The reference subclasses `arviz.InferenceData` and calls `az.summary(kind="stats")`.  arviz is an optional
dependency here, so the groups (`posterior`, `posterior_predictive`, `sample_stats`, `observed_data`) are small
dict-backed datasets and the summary statistics arviz would compute (mean, sd with ddof=1, shortest
highest-density interval, 3-decimal rounding) are restated in numpy.  `to_inference_data()` converts to a real
`az.InferenceData` when arviz is installed.

Multi-chain results (a capability the reference does not have) are pooled over chains wherever the reference
reads chain 0 (result_classes.py:243); with one chain the numbers coincide with the reference's definition.
"""
from __future__ import annotations

import pickle as pkl
from typing import Dict, List, Optional, Tuple

import numpy as np
import pandas as pd


class Coord:
    """Minimal stand-in for an xarray coordinate: `.values`, len(), iteration."""

    def __init__(self, values):
        self.values = np.asarray(list(values))

    def __len__(self):
        return len(self.values)

    def __iter__(self):
        return iter(self.values)

    def __repr__(self):
        return f"Coord({self.values!r})"


class LazyArray:
    """Read-only array that is computed on first use.

    The posterior-predictive quantities of a run (`concentration`, `prediction`: [draw, sample, cell_type] per chain)
    are generated on the device from the device-resident draws when somebody looks at them (`sccoda_predict`, K5)
    instead of being materialised on the host for every run (the reference builds them in per-draw python loops,
    scCODA_model.py:821-827).  Quacks like a numpy array: `shape`, `dtype`, `ndim`, `len()`, indexing, `np.asarray`,
    and every other ndarray attribute through the materialised value; `block(lo, hi)` fetches a range of draws
    without materialising the rest; pickles as the materialised array.

    fetch(lo, hi) -> ndarray with the draws [lo, hi) along `draw_axis`.
    """

    def __init__(self, shape, dtype, fetch, draw_axis: int = 0):
        self.shape = tuple(int(v) for v in shape)
        self.dtype = np.dtype(dtype)
        self._fetch = fetch
        self._draw_axis = int(draw_axis)
        self._value = None

    @property
    def ndim(self):
        return len(self.shape)

    @property
    def size(self):
        return int(np.prod(self.shape))

    @property
    def is_materialized(self):
        return self._value is not None

    def block(self, lo: int, hi: int) -> np.ndarray:
        """The draws [lo, hi) only."""
        if self._value is not None:
            idx = (slice(None),) * self._draw_axis + (slice(lo, hi),)
            return self._value[idx]
        return np.asarray(self._fetch(int(lo), int(hi)))

    def materialize(self) -> np.ndarray:
        if self._value is None:
            v = np.asarray(self._fetch(0, self.shape[self._draw_axis]))
            if v.shape != self.shape:
                raise ValueError(f"lazy array: fetched shape {v.shape}, expected {self.shape}")
            self._value = v
            self._fetch = None                     # drops the reference to the device-resident draws
        return self._value

    def with_leading_axis(self) -> "LazyArray":
        return LazyArray((1,) + self.shape, self.dtype, lambda lo, hi: self.block(lo, hi)[None], self._draw_axis + 1)

    def __array__(self, dtype=None, copy=None):
        v = self.materialize()
        return v.astype(dtype) if dtype is not None and np.dtype(dtype) != v.dtype else v

    def __getitem__(self, idx):
        return self.materialize()[idx]

    def __len__(self):
        return self.shape[0]

    def __getattr__(self, name):                    # only reached for attributes LazyArray itself does not have
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.materialize(), name)

    def __reduce__(self):
        return (np.asarray, (self.materialize(),))

    def __repr__(self):
        state = "materialized" if self._value is not None else "on demand"
        return f"LazyArray(shape={self.shape}, dtype={self.dtype}, {state})"


class Dataset:
    """Dict-backed group: variables are numpy arrays laid out [chain, draw, *dims] (or plain arrays for
    observed data); `ds["name"]`, `ds.name`, `ds.data_vars`, `ds.coords`, `ds.dims`."""

    def __init__(self, variables: Dict[str, np.ndarray], dims: Optional[Dict[str, List[str]]] = None,
                 coords: Optional[Dict[str, Coord]] = None):
        self._vars = dict(variables)
        self.dims = {k: list(v) for k, v in (dims or {}).items() if k in self._vars}
        self.coords = coords or {}

    @property
    def data_vars(self):
        return list(self._vars)

    def __getitem__(self, name):
        return self._vars[name]

    def __contains__(self, name):
        return name in self._vars

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        try:
            return self._vars[name]
        except KeyError:
            raise AttributeError(name) from None

    def keys(self):
        return self._vars.keys()

    def items(self):
        return self._vars.items()

    def __repr__(self):
        body = ", ".join(f"{k}{tuple(v.shape)}" for k, v in self._vars.items())
        return f"Dataset({body})"


def hdi(values: np.ndarray, hdi_prob: float = 0.94) -> Tuple[float, float]:
    """Shortest interval containing floor(hdi_prob * n) of the (non-NaN) sorted draws — arviz's unimodal HDI."""
    v = np.sort(values[~np.isnan(values)])
    n = v.size
    if n == 0:
        return np.nan, np.nan
    k = int(np.floor(hdi_prob * n))
    if k >= n:
        return float(v[0]), float(v[-1])
    widths = v[k:] - v[:n - k]
    i = int(np.argmin(widths))
    return float(v[i]), float(v[i + k])


def _stats_table(samples: np.ndarray, labels: List[str], hdi_prob: float, skipna: bool = False) -> pd.DataFrame:
    """az.summary(kind="stats") for a [draws, n_columns] sample: mean, sd (ddof=1), hdi bounds, rounded to 3 decimals."""
    lo_name = f"hdi_{100 * (1 - hdi_prob) / 2:g}%"
    hi_name = f"hdi_{100 * (1 - (1 - hdi_prob) / 2):g}%"
    rows = []
    for j in range(samples.shape[1]):
        col = samples[:, j]
        if skipna:
            col = col[~np.isnan(col)]
        if col.size == 0:
            rows.append((np.nan, np.nan, np.nan, np.nan))
            continue
        lo, hi = hdi(col, hdi_prob)
        sd = col.std(ddof=1) if col.size > 1 else np.nan
        rows.append((col.mean(), sd, lo, hi))
    df = pd.DataFrame(rows, columns=["mean", "sd", lo_name, hi_name], index=labels)
    return df.round(3)


class CAResult:
    """Compositional-analysis result.

    sampling_stats
        {"chain_length", "num_burnin", "acc_rate", "duration", "y_hat"} (scCODA_model.py:327-328)
    model_specs
        {"reference", "formula"}; "threshold_prob" is added by `summary_prepare` (result_classes.py:278)
    """

    def __init__(self, sampling_stats: dict, model_specs: dict, posterior: Dataset,
                 posterior_predictive: Optional[Dataset] = None, sample_stats: Optional[Dataset] = None,
                 observed_data: Optional[Dataset] = None):
        self.sampling_stats = sampling_stats
        self.model_specs = model_specs
        self.posterior = posterior
        self.posterior_predictive = posterior_predictive if posterior_predictive is not None else Dataset({})
        self.sample_stats = sample_stats if sample_stats is not None else Dataset({})
        self.observed_data = observed_data if observed_data is not None else Dataset({})
        self.is_sccoda = "ind" in self.posterior.data_vars                     # result_classes.py:98-101
        self.intercept_df, self.effect_df = self.summary_prepare()

    # ------------------------------------------------------------------ helpers
    def groups(self):
        return [g for g in ("posterior", "posterior_predictive", "sample_stats", "observed_data")
                if getattr(self, g).data_vars]

    def _pooled(self, name: str) -> np.ndarray:
        """[chain, draw, ...] -> [chain*draw, ...]"""
        a = np.asarray(self.posterior[name])
        return a.reshape((-1,) + a.shape[2:])

    # ------------------------------------------------------------------ summaries
    def summary_prepare(self, est_fdr: float = 0.05, hdi_prob: float = 0.94, **kwargs
                        ) -> Tuple[pd.DataFrame, pd.DataFrame]:
        """Intercept and effect DataFrames (result_classes.py:108-212).

        intercept_df: Final Parameter, HDI bounds, SD, Expected Sample — one row per cell type.
        effect_df:    Final Parameter (0 = not credible), HDI bounds, SD, Inclusion probability, Expected Sample,
                      log2-fold change — one row per (covariate, cell type).
        """
        cell_types = self.posterior.coords["cell_type"].values
        covariates = self.posterior.coords["covariate"].values
        K, D = len(cell_types), len(covariates)

        alpha = self._pooled("alpha")                                   # [S, K]
        beta = self._pooled("beta")                                     # [S, D, K]
        intercept_df = _stats_table(alpha, list(cell_types), hdi_prob)
        intercept_df.index = pd.Index(cell_types, name="Cell Type")
        effect_df = _stats_table(beta.reshape(beta.shape[0], D * K), list(range(D * K)), hdi_prob)
        effect_df.index = pd.MultiIndex.from_product([covariates, cell_types], names=["Covariate", "Cell Type"])
        hdis = [c for c in intercept_df.columns if "hdi" in c]

        intercept_df = self.complete_alpha_df(intercept_df)
        effect_df = self.complete_beta_df(intercept_df, effect_df, est_fdr)

        if self.is_sccoda:
            # credible interval of the effect given inclusion: b_raw * ind where ind >= 1e-3 (result_classes.py:175-196)
            ind = self._pooled("ind")
            b_raw = self._pooled("b_raw")
            sel = np.where(ind >= 1e-3, b_raw * ind, np.nan)            # [S, D, K-1]
            sel_stats = _stats_table(sel.reshape(sel.shape[0], -1), list(range(D * (K - 1))), hdi_prob, skipna=True)
            ref = self.model_specs["reference"]
            lo = np.zeros(D * K)
            hi = np.zeros(D * K)
            for d in range(D):
                nb = np.delete(np.arange(K), ref)
                lo[d * K + nb] = sel_stats[hdis[0]].to_numpy()[d * (K - 1):(d + 1) * (K - 1)]
                hi[d * K + nb] = sel_stats[hdis[1]].to_numpy()[d * (K - 1):(d + 1) * (K - 1)]
            effect_df[hdis[0]] = lo
            effect_df[hdis[1]] = hi

        nice = [h.replace("hdi_", "HDI ") for h in hdis]
        intercept_df = intercept_df.loc[:, ["final_parameter", hdis[0], hdis[1], "sd", "expected_sample"]].copy()
        intercept_df.columns = ["Final Parameter", nice[0], nice[1], "SD", "Expected Sample"]
        effect_df = effect_df.loc[:, ["final_parameter", hdis[0], hdis[1], "sd", "inclusion_prob", "expected_sample",
                                      "log_fold"]].copy()
        effect_df.columns = ["Final Parameter", nice[0], nice[1], "SD", "Inclusion probability", "Expected Sample",
                             "log2-fold change"]
        return intercept_df, effect_df

    def complete_alpha_df(self, intercept_df: pd.DataFrame) -> pd.DataFrame:
        """Expected sample for the intercepts (result_classes.py:314-342)."""
        intercept_df = intercept_df.rename(columns={"mean": "final_parameter"})
        y_bar = np.mean(np.sum(np.asarray(self.observed_data.y), axis=1))
        e = np.exp(intercept_df["final_parameter"].to_numpy())
        intercept_df["expected_sample"] = e / e.sum() * y_bar
        return intercept_df

    def complete_beta_df(self, intercept_df: pd.DataFrame, effect_df: pd.DataFrame, target_fdr: float = 0.05
                         ) -> pd.DataFrame:
        """Inclusion probability, FDR threshold, final effects, expected sample, log2 fold change
        (result_classes.py:214-312)."""
        beta = self._pooled("beta")                                      # [S, D, K]
        S, D, K = beta.shape
        nonzero = np.abs(beta) > 1e-3                                    # :249
        counts = nonzero.sum(axis=0)
        inc = (counts / S).reshape(-1)                                   # :250
        sums = np.where(nonzero, beta, 0.0).sum(axis=0)
        nz_mean = np.where(counts > 0, sums / np.maximum(counts, 1), 0.0).reshape(-1)   # :252-255
        effect_df["inclusion_prob"] = inc
        effect_df["mean_nonzero"] = nz_mean

        if self.is_sccoda:
            threshold = inclusion_threshold(inc, target_fdr)             # :262-276
            self.model_specs["threshold_prob"] = threshold
            effect_df["final_parameter"] = np.where(inc >= threshold, nz_mean, 0.0)     # :281-283
        else:
            effect_df["final_parameter"] = nz_mean

        y_bar = np.mean(np.sum(np.asarray(self.observed_data.y), axis=1))
        alpha_par = intercept_df["final_parameter"].to_numpy()
        a = np.exp(alpha_par)
        alpha_sample = a / a.sum() * y_bar
        final = effect_df["final_parameter"].to_numpy().reshape(D, K)
        b = np.exp(alpha_par[None, :] + final)
        b = b / b.sum(axis=1, keepdims=True) * y_bar
        effect_df["expected_sample"] = b.reshape(-1)
        effect_df["log_fold"] = np.log2(b / alpha_sample[None, :]).reshape(-1)
        return effect_df

    # ------------------------------------------------------------------ printing
    def summary(self, *args, **kwargs):
        """Short printed summary (result_classes.py:344-397)."""
        if args or kwargs:
            intercept_df, effect_df = self.summary_prepare(*args, **kwargs)
        else:
            intercept_df, effect_df = self.intercept_df, self.effect_df
        y_hat = self.sampling_stats.get("y_hat")
        dims = y_hat.shape if y_hat is not None else (10, 5)
        print("Compositional Analysis summary:")
        print("")
        print("Data: %d samples, %d cell types" % tuple(dims))
        print("Reference index: %s" % str(self.model_specs["reference"]))
        print("Formula: %s" % self.model_specs["formula"])
        print("")
        print("Intercepts:")
        print(intercept_df.loc[:, ["Final Parameter", "Expected Sample"]])
        print("")
        print("")
        print("Effects:")
        print(effect_df.loc[:, ["Final Parameter", "Expected Sample", "log2-fold change"]])

    def summary_extended(self, *args, **kwargs):
        """Extended printed summary with sampling diagnostics (result_classes.py:399-452)."""
        if args or kwargs:
            intercept_df, effect_df = self.summary_prepare(*args, **kwargs)
        else:
            intercept_df, effect_df = self.intercept_df, self.effect_df
        dims = self.sampling_stats["y_hat"].shape
        print("Compositional Analysis summary (extended):")
        print("")
        print("Data: %d samples, %d cell types" % tuple(dims))
        print("Reference index: %s" % str(self.model_specs["reference"]))
        print("Formula: %s" % self.model_specs["formula"])
        if self.is_sccoda:
            print("Spike-and-slab threshold: {threshold:.3f}".format(threshold=self.model_specs["threshold_prob"]))
        print("")
        print("MCMC Sampling: Sampled {num_results} chain states ({num_burnin} burnin samples) in {duration:.3f} sec. "
              "Acceptance rate: {ar:.1f}%".format(num_results=self.sampling_stats["chain_length"],
                                                  num_burnin=self.sampling_stats["num_burnin"],
                                                  duration=self.sampling_stats["duration"],
                                                  ar=(100 * self.sampling_stats["acc_rate"])))
        print("")
        print("Intercepts:")
        print(intercept_df)
        print("")
        print("")
        print("Effects:")
        print(effect_df)

    # ------------------------------------------------------------------ ground-truth comparisons
    def compare_parameters_to_truth(self, b_true: pd.Series, w_true: pd.Series, *args, **kwargs
                                    ) -> Tuple[pd.DataFrame, pd.DataFrame]:
        """Join the summaries with ground-truth intercepts / effects (result_classes.py:454-503)."""
        intercept_df, effect_df = self.summary_prepare(*args, **kwargs)
        intercept_df = intercept_df.rename(columns={"Final Parameter": "predicted"}).join(b_true.rename("truth"))
        effect_df = effect_df.rename(columns={"Final Parameter": "predicted"}).join(w_true.rename("truth"))
        for df in (intercept_df, effect_df):
            df["dist_to_truth"] = df["truth"] - df["predicted"]
            df["effect_correct"] = (df["truth"] == 0) == (df["predicted"] == 0)
        return intercept_df, effect_df

    def distance_to_truth(self) -> pd.DataFrame:
        """Absolute / relative error of the posterior-mode count matrix y_hat (result_classes.py:505-539)."""
        y = np.asarray(self.observed_data.y, dtype=np.float64)
        y_hat = self.sampling_stats["y_hat"]
        err = np.abs(y_hat - y)
        with np.errstate(divide="ignore", invalid="ignore"):
            rel = err / y
        rel[np.isinf(rel)] = 1.0
        rel[np.isnan(rel)] = 0.0
        out = pd.DataFrame({
            "Cell Type": ["Total"] + list(range(1, y.shape[1] + 1)),
            "Absolute Error": np.append(err.mean(), err.mean(axis=0)),
            "Relative Error": np.append(rel.mean(), rel.mean(axis=0)),
            "Actual Means": np.append(y.mean(), y.mean(axis=0)),
            "Predicted Means": np.append(y_hat.mean(), y_hat.mean(axis=0)),
        })
        return out

    # ------------------------------------------------------------------ decisions
    def credible_effects(self, est_fdr=None) -> pd.Series:
        """Boolean series: which effects are credible (result_classes.py:541-573)."""
        if type(est_fdr) == float:
            if est_fdr < 0 or est_fdr > 1:
                raise ValueError("est_fdr must be between 0 and 1!")
            _, eff_df = self.summary_prepare(est_fdr=est_fdr)
        else:
            eff_df = self.effect_df
        out = eff_df["Final Parameter"] != 0
        return out.rename("credible change")

    def set_fdr(self, est_fdr: float, *args, **kwargs):
        """Re-threshold the effects at another expected FDR (result_classes.py:594-619)."""
        self.intercept_df, self.effect_df = self.summary_prepare(est_fdr=est_fdr, *args, **kwargs)

    # ------------------------------------------------------------------ persistence / interop
    def save(self, path_to_file: str):
        """Pickle the result (result_classes.py:575-592)."""
        with open(path_to_file, "wb") as f:
            pkl.dump(self, file=f, protocol=4)

    def to_inference_data(self):
        """Convert to arviz.InferenceData (needs arviz)."""
        import arviz as az  # noqa: F401  (optional dependency)
        coords = {k: list(v.values) for k, v in self.posterior.coords.items()}
        kw = {}
        for group in ("posterior", "posterior_predictive", "sample_stats"):
            ds = getattr(self, group)
            if ds.data_vars:
                kw[group] = {k: np.asarray(v) for k, v in ds.items()}
        return az.from_dict(observed_data={k: np.asarray(v) for k, v in self.observed_data.items()}, coords=coords,
                            dims={k: v for k, v in self.posterior.dims.items()}, **kw)


def inclusion_threshold(inclusion_prob: np.ndarray, est_fdr: float) -> float:
    """Direct posterior probability approach (Newton et al. 2004) as in result_classes.py:262-276: the smallest
    inclusion probability c whose selected set {p >= c} has expected FDR mean(1 - p) below est_fdr, floored to
    3 decimals; 1.0 if there is none."""
    incs = inclusion_prob[inclusion_prob > 0]
    for c in np.unique(incs):
        if np.mean(1.0 - incs[incs >= c]) < est_fdr:
            return float(np.floor(c * 10 ** 3) / 10 ** 3)
    return 1.0
