"""
Design-matrix construction for `CompositionalAnalysis` (sccoda/util/comp_ana.py:73-75 uses patsy.dmatrix and
drops the intercept column).

patsy is an optional dependency: when it is importable it is used verbatim.  Otherwise a small treatment-coding
builder covers the formulas the reference's tests and tutorials use: sums of plain covariate names and
`C(name, Treatment('level'))` terms, e.g. "Condition", "Condition + Condition2",
"C(Condition, Treatment('Salm'))".  Interactions / transforms raise NotImplementedError (install patsy for those).
Column naming follows patsy: `name[T.level]` for the non-reference levels of a categorical covariate (levels
sorted; for pandas categoricals in category order), the bare name for numeric covariates.
"""
from __future__ import annotations

import re
from typing import List, Tuple

import numpy as np
import pandas as pd


def _split_terms(formula: str) -> List[str]:
    terms, depth, cur = [], 0, ""
    for ch in formula:
        if ch in "([":
            depth += 1
        elif ch in ")]":
            depth -= 1
        if ch == "+" and depth == 0:
            terms.append(cur.strip())
            cur = ""
        else:
            cur += ch
    terms.append(cur.strip())
    return [t for t in terms if t]


_C_TERM = re.compile(r"""^C\(\s*([A-Za-z_][\w.]*)\s*(?:,\s*Treatment\(\s*(?:reference\s*=\s*)?(['"])(.*?)\2\s*\)\s*)?\)$""")


def _levels(col: pd.Series) -> list:
    if isinstance(col.dtype, pd.CategoricalDtype):
        return list(col.cat.categories)
    return sorted(pd.unique(col.dropna()).tolist())


def _categorical(col: pd.Series, label: str, reference=None) -> Tuple[np.ndarray, List[str]]:
    levels = _levels(col)
    if reference is None:
        reference = levels[0]
    elif reference not in levels:
        raise ValueError(f"reference level {reference!r} is not a level of {label} (levels: {levels})")
    others = [lv for lv in levels if lv != reference]
    mat = np.stack([(col == lv).to_numpy(dtype=np.float64) for lv in others], axis=1) if others else \
        np.zeros((len(col), 0))
    return mat, [f"{label}[T.{lv}]" for lv in others]


def _builtin_design_matrix(formula: str, obs: pd.DataFrame) -> Tuple[np.ndarray, List[str]]:
    blocks, names = [], []
    for term in _split_terms(formula):
        if term in ("1", "+1"):
            continue
        if term in ("0", "-1"):
            raise NotImplementedError("formulas without intercept are not supported by the built-in builder")
        m = _C_TERM.match(term)
        if m:
            name, ref = m.group(1), m.group(3)
            if name not in obs.columns:
                raise KeyError(f"covariate {name!r} not found in obs")
            mat, nm = _categorical(obs[name], term, ref)
        elif re.fullmatch(r"[A-Za-z_][\w.]*", term):
            if term not in obs.columns:
                raise KeyError(f"covariate {term!r} not found in obs")
            col = obs[term]
            if pd.api.types.is_bool_dtype(col) or not pd.api.types.is_numeric_dtype(col):
                mat, nm = _categorical(col, term)
            else:
                mat, nm = col.to_numpy(dtype=np.float64)[:, None], [term]
        else:
            raise NotImplementedError(
                f"term {term!r}: the built-in design-matrix builder handles sums of covariate names and "
                f"C(name, Treatment('level')) terms only; install patsy for general formulas")
        blocks.append(mat)
        names.extend(nm)
    if not blocks:
        return np.zeros((len(obs), 0)), []
    return np.concatenate(blocks, axis=1), names


def design_matrix(formula: str, obs: pd.DataFrame) -> Tuple[np.ndarray, List[str]]:
    """Covariate matrix [N, D] WITHOUT the intercept column, and its column names."""
    try:
        import patsy  # type: ignore
    except ImportError:
        return _builtin_design_matrix(formula, obs)
    dm = patsy.dmatrix(formula, obs)
    return np.asarray(dm)[:, 1:], list(dm.design_info.column_names[1:])
