"""
Design-matrix construction for `CompositionalAnalysis` (sccoda/util/comp_ana.py:73-75 uses patsy.dmatrix and
drops the first column, normally the intercept).

patsy is an optional dependency: when it is importable it is used verbatim.  Otherwise the builder below covers the
formula language the reference's users write, with patsy's conventions:

  terms      `a + b`, interactions `a:b`, `a*b` (= a + b + a:b), `(a + b)**2`, removal `- a`, intercept `1` / `0` / `-1`
  factors    a covariate name; `C(name)`, `C(name, Treatment('level'))` / `C(name, Treatment(reference='level'))`;
             numeric transforms `np.log(x)`, `np.log1p(x)`, `np.sqrt(x)`, `np.exp(x)`, ...; `I(<arithmetic>)`;
             the stateful transforms `center(x)`, `standardize(x)` / `scale(x)`
  coding     treatment (dummy) coding of categorical factors, reference = first level (levels sorted; category order for
             pandas categoricals); strings / booleans are categorical, numbers numeric
  columns    `name[T.level]`, numeric factors by their code, interactions joined with `:`; the left-most factor of an
             interaction varies fastest; terms without numeric factors come first, then by degree (patsy's ordering)
  redundancy  a categorical factor is treatment-coded inside a term when the term without it is in the model too, and
             coded with ALL its levels (`name[level]`) otherwise — e.g. `0 + Condition`, `Condition:dose` — as patsy
             does; without an intercept the reference still drops the first column (comp_ana.py:75) — reproduced

Not covered (raises NotImplementedError, install patsy): interactions of two or more categorical factors whose
lower-order terms are missing (patsy switches to a mixed full-rank coding there), contrasts other than Treatment,
`a/b`, `a %in% b`.
"""
from __future__ import annotations

import itertools
import re
from typing import Dict, List, Tuple

import numpy as np
import pandas as pd

_NAME = r"[A-Za-z_][\w.]*"
_C_TERM = re.compile(r"""^C\(\s*(""" + _NAME + r""")\s*(?:,\s*Treatment\(\s*(?:reference\s*=\s*)?(['"])(.*?)\2\s*\)\s*)?\)$""")
_CALL = re.compile(r"^(np\.\w+|numpy\.\w+|I|center|standardize|scale)\((.*)\)$", re.S)


# ------------------------------------------------------------------------------------------------------------
# formula -> list of terms (each a tuple of factor codes); the empty tuple is the intercept
# ------------------------------------------------------------------------------------------------------------
def _tokens(formula: str) -> List[str]:
    """Split at top level into factors and the operators + - : * ** ( )."""
    out, cur, depth, i = [], "", 0, 0
    f = formula.strip()
    while i < len(f):
        ch = f[i]
        if depth == 0 and cur.strip() == "" and ch == "(":
            out.append("(")                                    # grouping parenthesis (no function name before it)
            i += 1
            continue
        if ch in "([":
            depth += 1
        elif ch in ")]":
            if depth == 0:
                if cur.strip():
                    out.append(cur.strip())
                out.append(")")
                cur = ""
                i += 1
                continue
            depth -= 1
        if depth == 0 and ch in "+-:*" and not (ch == "-" and cur.strip().endswith(("e", "E")) and cur.strip()[:-1].replace(".", "").isdigit()):
            if cur.strip():
                out.append(cur.strip())
            cur = ""
            if ch == "*" and f[i:i + 2] == "**":
                out.append("**")
                i += 2
                continue
            out.append(ch)
        else:
            cur += ch
        i += 1
    if cur.strip():
        out.append(cur.strip())
    return out


class _Parser:
    """Recursive descent over patsy's operator precedence: ** > : > * > + -.  A value is an ordered list of terms."""

    def __init__(self, toks):
        self.t, self.i = toks, 0

    def peek(self):
        return self.t[self.i] if self.i < len(self.t) else None

    def take(self):
        tok = self.peek()
        self.i += 1
        return tok

    @staticmethod
    def _uniq(terms):
        seen, out = set(), []
        for t in terms:
            key = frozenset(t)
            if key not in seen:
                seen.add(key)
                out.append(t)
        return out

    @staticmethod
    def _interact(a, b):
        return _Parser._uniq([tuple(dict.fromkeys(x + y)) for x in a for y in b])

    def expr(self):
        terms, removed, first = [], [], True
        while self.peek() is not None and self.peek() != ")":
            op = "+"
            if self.peek() in "+-":
                op = self.take()
            elif not first:
                raise ValueError(f"unexpected token {self.peek()!r} in formula")
            first = False
            val = self.product()
            if op == "+":
                terms = self._uniq(terms + val)
            else:
                removed += val
        rm = {frozenset(t) for t in removed}
        return [t for t in terms if frozenset(t) not in rm], removed

    def product(self):
        val = self.interaction()
        while self.peek() == "*":
            self.take()
            rhs = self.interaction()
            val = self._uniq(val + rhs + self._interact(val, rhs))
        return val

    def interaction(self):
        val = self.power()
        while self.peek() == ":":
            self.take()
            val = self._interact(val, self.power())
        return val

    def power(self):
        val = self.atom()
        if self.peek() == "**":
            self.take()
            n = int(self.take())
            out = list(val)
            for _ in range(n - 1):
                out = self._uniq(out + self._interact(out, val))
            val = out
        return val

    def atom(self):
        tok = self.take()
        if tok is None:
            raise ValueError("formula ends unexpectedly")
        if tok == "(":
            terms, removed = self.expr()
            if self.take() != ")":
                raise ValueError("unbalanced parenthesis in formula")
            return terms
        if tok in ("1",):
            return [()]
        if tok in ("0",):
            return [("__no_intercept__",)]
        return [(tok,)]


def _parse(formula: str):
    """-> (terms without the intercept, has_intercept)"""
    rhs = formula.split("~", 1)[-1]
    p = _Parser(_tokens(rhs))
    terms, removed = p.expr()
    if p.peek() is not None:
        raise ValueError(f"unexpected token {p.peek()!r} in formula")
    intercept = True
    if any(t == ("__no_intercept__",) for t in terms) or any(t == () for t in removed):
        intercept = False
    terms = [t for t in terms if t not in ((), ("__no_intercept__",))]
    return terms, intercept


# ------------------------------------------------------------------------------------------------------------
# factors
# ------------------------------------------------------------------------------------------------------------
def _levels(col: pd.Series) -> list:
    if isinstance(col.dtype, pd.CategoricalDtype):
        return list(col.cat.categories)
    return sorted(pd.unique(col.dropna()).tolist())


def _is_categorical(col: pd.Series) -> bool:
    return pd.api.types.is_bool_dtype(col) or not pd.api.types.is_numeric_dtype(col)


class _Factor:
    """One factor of the formula evaluated on obs: numeric (one column) or categorical (levels + reference)."""

    def __init__(self, code: str, obs: pd.DataFrame):
        self.code = code
        self.numeric = True
        self.levels, self.reference, self.col = None, None, None
        m = _C_TERM.match(code)
        if m:
            name, ref = m.group(1), m.group(3)
            self._categorical(self._column(name, obs), ref)
            return
        if re.fullmatch(_NAME, code):
            col = self._column(code, obs)
            if _is_categorical(col):
                self._categorical(col, None)
            else:
                self.values = col.to_numpy(dtype=np.float64)
            return
        m = _CALL.match(code)
        if m:
            self.values = self._evaluate(m.group(1), m.group(2), obs)
            return
        if code.startswith("C("):
            raise NotImplementedError(f"factor {code!r}: only Treatment contrasts are built in; install patsy for others")
        raise NotImplementedError(f"factor {code!r} is not understood by the built-in design-matrix builder; install patsy")

    @staticmethod
    def _column(name: str, obs: pd.DataFrame) -> pd.Series:
        if name not in obs.columns:
            raise KeyError(f"covariate {name!r} not found in obs")
        return obs[name]

    def _categorical(self, col: pd.Series, ref):
        self.numeric, self.col = False, col
        self.levels = _levels(col)
        if ref is None:
            ref = self.levels[0]
        elif ref not in self.levels:
            # patsy compares the quoted level with the data: try the level's own type ('1' for an integer level 1)
            alt = [lv for lv in self.levels if str(lv) == ref]
            if not alt:
                raise ValueError(f"reference level {ref!r} is not a level of {self.code} (levels: {self.levels})")
            ref = alt[0]
        self.reference = ref

    @staticmethod
    def _evaluate(fn: str, arg: str, obs: pd.DataFrame) -> np.ndarray:
        env: Dict[str, object] = {"np": np, "numpy": np}
        for name in obs.columns:
            if isinstance(name, str) and name.isidentifier() and not _is_categorical(obs[name]):
                env[name] = obs[name].to_numpy(dtype=np.float64)
        try:
            inner = eval(arg, {"__builtins__": {}}, env)       # arithmetic on the numeric covariates only
        except Exception as exc:
            raise ValueError(f"cannot evaluate {arg!r} in the formula: {exc}") from exc
        v = np.asarray(inner, dtype=np.float64)
        if fn == "I":
            out = v
        elif fn == "center":
            out = v - v.mean()
        elif fn in ("standardize", "scale"):
            out = (v - v.mean()) / v.std(ddof=0)              # patsy: ddof = 0
        else:
            out = np.asarray(eval(fn, {"__builtins__": {}}, env)(v), dtype=np.float64)
        if out.ndim == 0:
            out = np.full(len(obs), float(out))
        return out

    def columns(self, full_rank: bool = False) -> Tuple[np.ndarray, List[str]]:
        if self.numeric:
            return self.values[:, None], [self.code]
        lv = self.levels if full_rank else [x for x in self.levels if x != self.reference]
        mat = np.stack([(self.col == x).to_numpy(dtype=np.float64) for x in lv], axis=1) if lv else np.zeros((len(self.col), 0))
        return mat, [f"{self.code}[{x}]" if full_rank else f"{self.code}[T.{x}]" for x in lv]


def _builtin_design_matrix(formula: str, obs: pd.DataFrame) -> Tuple[np.ndarray, List[str]]:
    terms, intercept = _parse(formula)
    factors: Dict[str, _Factor] = {}
    for t in terms:
        for code in t:
            if code not in factors:
                factors[code] = _Factor(code, obs)
    # patsy's term order: grouped by the set of numeric factors a term contains (none first, then in order of first
    # appearance), inside a group by degree; ties keep the formula's order
    groups: List[frozenset] = [frozenset()]
    for t in terms:
        g = frozenset(c for c in t if factors[c].numeric)
        if g not in groups:
            groups.append(g)
    order = sorted(range(len(terms)), key=lambda i: (groups.index(frozenset(c for c in terms[i] if factors[c].numeric)),
                                                     len(terms[i]), i))
    terms = [terms[i] for i in order]
    present = {frozenset(t) for t in terms}
    if intercept:
        present.add(frozenset())
    blocks, names = [], []
    for t in terms:
        # patsy's redundancy rule: a categorical factor is treatment-coded inside a term when the term without that
        # factor is part of the model as well (its span is then already covered), and coded with all its levels otherwise
        full = {c: (not factors[c].numeric) and frozenset(x for x in t if x != c) not in present for c in t}
        if sum(full.values()) > 0 and sum(1 for c in t if not factors[c].numeric) > 1:
            raise NotImplementedError(
                f"interaction {':'.join(t)!r} of categorical factors without its lower-order terms: patsy codes such terms "
                f"with a mixed full-rank scheme that the built-in builder does not reproduce; install patsy")
        cols = [factors[c].columns(full_rank=full[c]) for c in t]
        for c in t:
            if full[c]:
                present.add(frozenset(x for x in t if x != c))       # all levels of c: the term without c is spanned now
        sizes = [m.shape[1] for m, _ in cols]
        # the left-most factor varies fastest (patsy: "for consistency with R")
        for combo in itertools.product(*[range(n) for n in reversed(sizes)]):
            combo = combo[::-1]
            v = np.ones(len(obs))
            for (m, _), j in zip(cols, combo):
                v = v * m[:, j]
            blocks.append(v[:, None])
            names.append(":".join(nm[j] for (_, nm), j in zip(cols, combo)))
    mat = np.concatenate(blocks, axis=1) if blocks else np.zeros((len(obs), 0))
    if not intercept:
        # the reference drops the first column of patsy's matrix whatever it is (comp_ana.py:75)
        return mat[:, 1:], names[1:]
    return mat, names


def design_matrix(formula: str, obs: pd.DataFrame) -> Tuple[np.ndarray, List[str]]:
    """Covariate matrix [N, D] WITHOUT the first (intercept) column, and its column names."""
    try:
        import patsy  # type: ignore
    except ImportError:
        return _builtin_design_matrix(formula, obs)
    dm = patsy.dmatrix(formula, obs)
    return np.asarray(dm)[:, 1:], list(dm.design_info.column_names[1:])
