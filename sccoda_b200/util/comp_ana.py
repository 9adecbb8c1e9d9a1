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
