"""Design-matrix builder with mgcv's *structure* (host side, numpy/scipy only).

The reference builds X_fe / X_re / S per SDE parameter with ``mgcv::gam(..., fit = FALSE)`` and
stacks them block-diagonally (R/sde.R:378-455, ``make_mat``; ``bdiag_check`` R/utility.R:13-28).
mgcv is an R package and is not available here, so this module produces matrices with the same
layout contract:

  * rows ``j*n .. (j+1)*n-1`` of X_fe and X_re belong to SDE parameter ``j``;
  * the first ``nsdf`` columns of a parameter's model matrix (intercept + linear terms) go to
    X_fe, smooth columns go to X_re (R/sde.R:412-421);
  * one penalty block per smooth, ``ncol_re = sapply(S, ncol)`` (R/sde.R:431-433);
  * ``s(x, k = K)`` has K-1 columns after the sum-to-zero constraint is absorbed
    (tests/testthat/test_sde.R:68-71 pins 4 columns for k = 5) and ``s(ID, bs = "re")`` has one
    column per level with S = I.

The smooth basis itself is a cubic B-spline on equally spaced knots with a second-difference
(P-spline) penalty plus a small ridge so that S is full rank, which ``nllk_sde`` needs
(src/nllk/nllk_sde.hpp:109-111; the reference sticks to shrinkage bases "cs"/"ts" for the same
reason).  It is not numerically identical to mgcv's "cs" basis; the likelihood engine only sees
(X_fe, X_re, S, ncol_re) so this does not matter for parity of the objective.
"""
from __future__ import annotations

import re as _re
from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp


# --------------------------------------------------------------------------------------------
# basis functions
# --------------------------------------------------------------------------------------------
def bspline_basis(x, k, lo=None, hi=None, degree=3):
    """Dense [n, k] B-spline basis on equally spaced knots covering [lo, hi]."""
    x = np.asarray(x, dtype=float)
    lo = float(np.min(x)) if lo is None else float(lo)
    hi = float(np.max(x)) if hi is None else float(hi)
    if hi <= lo:
        hi = lo + 1.0
    nseg = k - degree
    if nseg < 1:
        raise ValueError("k must be at least degree + 1")
    h = (hi - lo) / nseg
    # scaled coordinate; basis j is the cardinal cubic B-spline centred on knot j - 1... use
    # the closed form of the uniform cubic B-spline pieces
    t = (x - lo) / h
    seg = np.clip(np.floor(t).astype(np.int64), 0, nseg - 1)
    u = t - seg
    if degree != 3:
        raise NotImplementedError("cubic only")
    w0 = (1 - u) ** 3 / 6.0
    w1 = (3 * u ** 3 - 6 * u ** 2 + 4) / 6.0
    w2 = (-3 * u ** 3 + 3 * u ** 2 + 3 * u + 1) / 6.0
    w3 = u ** 3 / 6.0
    B = np.zeros((x.size, k))
    rows = np.arange(x.size)
    for j, w in enumerate((w0, w1, w2, w3)):
        B[rows, seg + j] = w
    return B


def sum_to_zero_transform(colmeans):
    """Null-space basis Z [k, k-1] of the constraint 1'X b = 0 (mgcv absorbs it by QR)."""
    c = np.asarray(colmeans, dtype=float).reshape(-1, 1)
    q, _ = np.linalg.qr(c, mode="complete")
    return q[:, 1:]


def second_diff_penalty(k):
    D = np.diff(np.eye(k), n=2, axis=0)
    return D.T @ D


@dataclass
class SmoothSpec:
    """Frozen description of one ``s(var, k=K)`` term so it can be re-evaluated on new data."""
    var: str
    k: int
    lo: float
    hi: float
    Z: np.ndarray
    S: np.ndarray

    def basis(self, x):
        return bspline_basis(x, self.k, self.lo, self.hi) @ self.Z


def make_smooth(var, x, k=10, ridge=1e-2):
    x = np.asarray(x, dtype=float)
    lo, hi = float(np.min(x)), float(np.max(x))
    B = bspline_basis(x, k, lo, hi)
    Z = sum_to_zero_transform(B.mean(axis=0))
    S = Z.T @ second_diff_penalty(k) @ Z + ridge * np.eye(k - 1)
    S = 0.5 * (S + S.T)
    return SmoothSpec(var, k, lo, hi, Z, S)


# --------------------------------------------------------------------------------------------
# formula mini-language:  "~ 1", "~ x1 + s(time, k = 10, bs = 'cs') + s(ID, bs = 're')"
# --------------------------------------------------------------------------------------------
@dataclass
class Term:
    kind: str            # "linear" | "smooth" | "re"
    var: str
    k: int = 10


def parse_formula(formula):
    f = formula.strip()
    if f.startswith("~"):
        f = f[1:]
    terms = []
    depth = 0
    cur = ""
    parts = []
    for ch in f:
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        if ch == "+" and depth == 0:
            parts.append(cur)
            cur = ""
        else:
            cur += ch
    parts.append(cur)
    for p in parts:
        p = p.strip()
        if p in ("", "1"):
            continue
        m = _re.fullmatch(r"s\((.*)\)", p)
        if m:
            args = [a.strip() for a in m.group(1).split(",")]
            var = args[0]
            k = 10
            bs = "tp"
            for a in args[1:]:
                key, _, val = a.partition("=")
                key, val = key.strip(), val.strip().strip("\"'")
                if key == "k":
                    k = int(val)
                elif key == "bs":
                    bs = val
            terms.append(Term("re", var) if bs == "re" else Term("smooth", var, k))
        else:
            terms.append(Term("linear", p))
    return terms


@dataclass
class Design:
    X_fe: sp.csr_matrix
    X_re: sp.csr_matrix
    S: sp.csr_matrix
    ncol_fe: np.ndarray
    ncol_re: np.ndarray
    names_fe: list
    names_re: list
    names_ncol_re: list
    specs: list = field(default_factory=list)    # per parameter: list of (Term, spec)


def make_design(formulas, data, n=None):
    """formulas: ordered dict {par_name: "~ ..."}; data: mapping of column name -> array.

    Mirrors SDE$make_mat (R/sde.R:378-455).  Returns a :class:`Design`.
    """
    if n is None:
        n = len(next(iter(data.values())))
    X_list_fe, X_list_re, S_list = [], [], []
    ncol_fe, ncol_re, names_fe, names_re, names_ncol_re, specs = [], [], [], [], [], []
    for par_name, form in formulas.items():
        terms = parse_formula(form)
        fe_cols = [np.ones(n)]
        fe_names = [par_name + ".(Intercept)"]
        re_blocks = []
        par_specs = []
        for t in terms:
            if t.var not in data:
                raise KeyError(f"covariate '{t.var}' not found in data")   # R/sde.R:103-108
            if t.kind == "linear":
                fe_cols.append(np.asarray(data[t.var], dtype=float))
                fe_names.append(f"{par_name}.{t.var}")
                par_specs.append((t, None))
            elif t.kind == "smooth":
                spec = make_smooth(t.var, data[t.var], t.k)
                Xb = sp.csr_matrix(spec.basis(data[t.var]))
                re_blocks.append(Xb)
                S_list.append(sp.csr_matrix(spec.S))
                ncol_re.append(t.k - 1)
                names_re += [f"{par_name}.s({t.var}).{i + 1}" for i in range(t.k - 1)]
                names_ncol_re.append(f"{par_name}.s({t.var})")
                par_specs.append((t, spec))
            elif t.kind == "re":
                codes, levels = factor_codes(data[t.var])
                L = len(levels)
                Xb = sp.csr_matrix((np.ones(n), (np.arange(n), codes)), shape=(n, L))
                re_blocks.append(Xb)
                S_list.append(sp.identity(L, format="csr"))
                ncol_re.append(L)
                names_re += [f"{par_name}.s({t.var}).{i + 1}" for i in range(L)]
                names_ncol_re.append(f"{par_name}.s({t.var})")
                par_specs.append((t, levels))
        X_list_fe.append(sp.csr_matrix(np.column_stack(fe_cols)))
        names_fe += fe_names
        ncol_fe.append(len(fe_cols))
        X_list_re.append(sp.hstack(re_blocks, format="csr") if re_blocks
                         else sp.csr_matrix((n, 0)))
        specs.append(par_specs)
    X_fe = sp.block_diag(X_list_fe, format="csr")
    X_re = sp.block_diag(X_list_re, format="csr")
    S = sp.block_diag(S_list, format="csr") if S_list else None
    return Design(X_fe, X_re, S, np.asarray(ncol_fe), np.asarray(ncol_re, dtype=np.int64),
                  names_fe, names_re, names_ncol_re, specs)


def factor_codes(x):
    """R factor codes (0-based) in order of first appearance, plus the level list."""
    x = np.asarray(x)
    levels, first = np.unique(x, return_index=True)
    order = np.argsort(first)
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    codes = rank[np.searchsorted(levels, x)]
    return codes.astype(np.int64), list(levels[order])
