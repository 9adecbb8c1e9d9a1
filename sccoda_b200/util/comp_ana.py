"""
Model factory — same call as the reference's `sccoda.util.comp_ana.CompositionalAnalysis`
(sccoda/util/comp_ana.py:14-130): builds the count and covariate matrices from an AnnData-like object and
returns a `scCODAModel` whose samplers run on the B200 engine.
"""
from __future__ import annotations

from typing import Union

import numpy as np

from ..model import scCODA_model as dm
from .formula import design_matrix


class CompositionalAnalysis:
    """`model = CompositionalAnalysis(data, formula="covariate1 + covariate2", reference_cell_type="CellTypeA")`

    data
        AnnData (or `CompositionData`): counts in `.X`, covariates in `.obs`, cell types in `.var.index`.
    formula
        R-style formula for the covariate matrix (see util/formula.py; patsy is used when installed).
    reference_cell_type
        cell type name, column index, or "automatic": the cell type with the smallest dispersion of relative
        abundance among those that are zero in less than `automatic_reference_absence_threshold` of the samples.
    """

    def __new__(cls, data, formula: str, reference_cell_type: Union[str, int] = "automatic",
                automatic_reference_absence_threshold: float = 0.05) -> "dm.scCODAModel":
        cell_types = data.var.index.to_list()
        counts = np.asarray(data.X).astype("float64")                       # comp_ana.py:70 (a copy)
        covariates, covariate_names = design_matrix(formula, data.obs)      # comp_ana.py:73-75

        if isinstance(reference_cell_type, str) and reference_cell_type == "automatic":
            ref_index = automatic_reference(counts, automatic_reference_absence_threshold)
            print(f"Automatic reference selection! Reference cell type set to {cell_types[ref_index]}")
        elif reference_cell_type in cell_types:
            ref_index = cell_types.index(reference_cell_type)
        elif isinstance(reference_cell_type, (int, np.integer)) and 0 <= reference_cell_type < len(cell_types):
            ref_index = int(reference_cell_type)
        else:
            raise NameError("Reference index is not a valid cell type name or numerical index!")

        return dm.scCODAModel(reference_cell_type=ref_index, covariate_matrix=np.array(covariates),
                              data_matrix=counts, cell_types=cell_types, covariate_names=covariate_names,
                              formula=formula)


def automatic_reference(counts: np.ndarray, absence_threshold: float = 0.05) -> int:
    """Dispersion-based reference selection (comp_ana.py:79-91)."""
    zero_share = np.mean(counts == 0, axis=0)
    candidates = np.flatnonzero(zero_share < absence_threshold)
    if candidates.size == 0:
        raise ValueError("No cell types that have large enough presence! "
                         "Please increase automatic_reference_absence_threshold")
    rel = counts / counts.sum(axis=1, keepdims=True)
    dispersion = rel.var(axis=0) / rel.mean(axis=0)
    best = dispersion[candidates].min()
    return int(np.flatnonzero(dispersion == best)[0])


def automatic_reference_batch(counts: np.ndarray, absence_threshold: float = 0.05) -> np.ndarray:
    """`automatic_reference` for a batch of datasets at once (counts [B,N,K] -> reference index [B]): the rule of
    comp_ana.py:79-91 vectorised, for sweeps where every replicate picks its own reference."""
    counts = np.asarray(counts, dtype=np.float64)
    zero_share = np.mean(counts == 0, axis=1)                               # [B,K]
    ok = zero_share < absence_threshold
    if not np.all(ok.any(axis=1)):
        raise ValueError("No cell types that have large enough presence! "
                         "Please increase automatic_reference_absence_threshold")
    rel = counts / counts.sum(axis=2, keepdims=True)
    dispersion = rel.var(axis=1) / rel.mean(axis=1)                         # [B,K]
    best = np.where(ok, dispersion, np.inf).min(axis=1, keepdims=True)
    return np.argmax(dispersion == best, axis=1).astype(np.int32)           # first cell type with that dispersion


def reference_sweep(data, formula: str, num_results: int = int(20e3), num_burnin: int = int(5e3), num_chains: int = 4,
                    est_fdr: float = 0.05, seed: int = 0, references=None):
    """Refit the model once per reference cell type — all refits in ONE launch — and report which effects are
    credible under how many references.

    The reference's tutorial does this with a python loop of K sequential `sample_hmc` runs
    (tutorials/Modeling_options_and_result_analysis.ipynb, "reference cell type" section: 8 refits, 751 s); here the
    K refits are K datasets of one batch (same counts, different reference index), `num_chains` chains each.
    Returns (times_credible DataFrame [covariate x cell type], results dict of the sweep incl. `inclusion_prob`,
    `final_effect`, `threshold` per refit).  BASELINE.json configs[3]."""
    import pandas as pd
    from .. import _native as nat
    from .. import scheduler as sch
    from .._pack import apply_pseudocount, pack
    cell_types = data.var.index.to_list()
    counts = apply_pseudocount(np.asarray(data.X).astype("float64"))        # scCODA_model.py:94-96
    covariates, covariate_names = design_matrix(formula, data.obs)
    x = np.asarray(covariates, dtype=np.float64)
    K = len(cell_types)
    refs = np.arange(K) if references is None else np.asarray(
        [cell_types.index(r) if r in cell_types else int(r) for r in references])
    R = len(refs)
    packed = pack(np.broadcast_to(x, (R,) + x.shape), np.broadcast_to(counts, (R,) + counts.shape), refs,
                  dtype=np.float32, pseudocount=False)
    cfg = sch.SweepConfig(method=nat.METHOD_HMC, num_results=num_results, num_burnin=num_burnin, chains=num_chains,
                          seed=seed, est_fdr=est_fdr)
    shard = sch.sample_shard(None, None, None, cfg, 0, R, packed=packed)
    res = {k: v.cpu().numpy() for k, v in sch.shard_results(shard, cfg).items()}
    res["threshold"], res["final_effect"] = sch.credible_effect_calls(res["inclusion_prob"], res["mean_nonzero"], est_fdr)
    res["references"] = [cell_types[r] for r in refs]
    res["kernel_ms"] = shard.out.kernel_ms
    D = x.shape[1]
    times = (res["final_effect"] != 0).sum(axis=0).reshape(D, K)
    times_credible = pd.DataFrame(times, index=pd.Index(covariate_names, name="Covariate"),
                                  columns=pd.Index(cell_types, name="Cell Type"))
    return times_credible, res
