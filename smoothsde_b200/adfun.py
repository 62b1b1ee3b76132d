"""`MakeADFun`-shaped objective on top of the CUDA engine.

In the reference, ``SDE$setup()`` calls ``TMB::MakeADFun(data, parameters, map, random)`` twice
(R/sde.R:656-669) and then only uses the returned closures: ``obj$par``, ``obj$fn``, ``obj$gr``
(R/sde.R:694-697), ``obj$report()`` (R/sde.R:1324) and ``obj$env$last.par.best`` (sdreport).
:class:`ADFun` provides the same members for the parameter lists the hot path knows:

    CTCRW : log_sigma_obs, coeff_fe, log_lambda, coeff_re      (nllk_ctcrw.hpp:135-140)
    BM/OU : coeff_fe, log_lambda, [log_decay], coeff_re        (nllk_sde.hpp:42-45)

``map`` follows TMB: per parameter a vector of factor codes, ``None``/NaN fixes an entry at its
initial value, equal codes tie entries together (``fixpar`` becomes ``map$coeff_fe`` with NA at
the fixed positions, R/sde.R:621-632).  ``random="coeff_re"`` asks for the Laplace-marginal
objective; it is provided by :mod:`smoothsde_b200.laplace`.
"""
from __future__ import annotations

import numpy as np

from ._lib import EngineError
from .engine import Engine

PAR_ORDER = ("log_sigma_obs", "coeff_fe", "log_lambda", "log_decay", "coeff_re")


class Env:
    """The part of TMB's ``obj$env`` that R/sde.R and sdreport read."""

    def __init__(self, last_par):
        self.last_par = last_par.copy()
        self.last_par_best = last_par.copy()
        self.value_best = np.inf
        self.random = None


class ADFun:
    def __init__(self, data, parameters, map=None, random=None, device=0, engine=None, laplace_opts=None):
        self.data = data
        self.engine = engine if engine is not None else Engine.from_data(data, device=device)
        layout = self.engine.layout                      # name -> (offset, size) in the full vector
        map = dict(map or {})
        # ---- full parameter vector in the templates' PARAMETER order
        full = np.zeros(self.engine.n_par)
        names = np.empty(self.engine.n_par, dtype=object)
        for nm in PAR_ORDER:
            if nm == "log_decay" and nm not in layout:
                # exists in the list for BM/OU (R/sde.R:504-507) but is mapped off without decay
                # terms (R/sde.R:648): it is then absent from the engine's vector
                if nm in parameters and nm in map and not _all_na(map[nm]):
                    raise ValueError("log_decay is free but the data list has no decay term (t_decay)")
                continue
            if nm not in layout:
                continue
            off, size = layout[nm]
            val = np.atleast_1d(np.asarray(parameters.get(nm, np.zeros(size)), dtype=float))
            if val.size != size:
                raise ValueError(f"parameter '{nm}' has length {val.size}, the data imply {size}")
            full[off:off + size] = val
            names[off:off + size] = nm
        # ---- map: group id per entry (-1 = fixed)
        group = np.arange(full.size)
        for nm, codes in map.items():
            if nm not in layout:
                continue
            off, size = layout[nm]
            codes = np.atleast_1d(np.asarray(codes, dtype=object))
            if codes.size != size:
                raise ValueError(f"map${nm} has length {codes.size}, expected {size}")
            first = {}
            for i, c in enumerate(codes):
                if c is None or (isinstance(c, float) and np.isnan(c)):
                    group[off + i] = -1
                else:
                    group[off + i] = first.setdefault(c, off + i)
        self._full0 = full
        self._group = group
        self.names_full = names
        self.random = random
        if random is not None:
            if random != "coeff_re":
                raise ValueError("only random = 'coeff_re' exists in smoothSDE (R/sde.R:522-524)")
            off, size = layout["coeff_re"]
            self._is_random = np.zeros(full.size, dtype=bool)
            self._is_random[off:off + size] = True
        else:
            self._is_random = np.zeros(full.size, dtype=bool)
        free = np.unique(group[group >= 0])
        self._free = free                                   # representative index of every free group
        self._active = free[~self._is_random[free]]         # what optim sees (obj$par)
        self._rand = free[self._is_random[free]]
        self.par = full[self._active].copy()
        self.names = names[self._active]
        self.env = Env(full)
        self.env.random = self._rand
        self._laplace = None
        if random is not None and self._rand.size:
            from .laplace import Laplace
            self._laplace = Laplace(self, **(laplace_opts or {}))

    # ------------------------------------------------------------------------------------
    def full_from(self, x, b=None):
        """Scatter active (and optionally random) values into the full parameter vector."""
        p = self._full0.copy()
        vals = np.empty(p.size)
        vals[:] = np.nan
        vals[self._active] = np.asarray(x, dtype=float)
        if self._rand.size:
            vals[self._rand] = self.env.last_par[self._rand] if b is None else np.asarray(b, dtype=float)
        ok = self._group >= 0
        p[ok] = vals[self._group[ok]]
        return p

    def reduce_grad(self, g_full, idx):
        """Sum the full-vector gradient over tied entries and pick the groups in `idx`."""
        out = np.zeros(self._full0.size)
        ok = self._group >= 0
        np.add.at(out, self._group[ok], g_full[ok])
        return out[idx]

    def joint(self, p_full, order=1):
        try:
            v, g = self.engine.eval(p_full, order=order)
        except EngineError as e:
            # SSDE_ERR_NUMERIC (F <= 0 in the filter, ...) at these parameters: the objective is undefined,
            # an optimiser must see Inf / NaN and step back, not an exception
            if e.code != 5:
                raise
            v, g = np.inf, (np.full(np.asarray(p_full).size, np.nan) if order >= 1 else None)
        self.env.last_par = p_full.copy()
        return v, g

    def _remember(self, v, p_full):
        if np.isfinite(v) and v < self.env.value_best:
            self.env.value_best = v
            self.env.last_par_best = p_full.copy()

    # ------------------------------------------------------------------------------------
    def fn(self, x=None):
        x = self.par if x is None else x
        if self._laplace is not None:
            return self._laplace.fn(np.asarray(x, dtype=float))
        p = self.full_from(x)
        v, _ = self.joint(p, order=0)
        self._remember(v, p)
        return v

    def gr(self, x=None):
        x = self.par if x is None else x
        if self._laplace is not None:
            return self._laplace.gr(np.asarray(x, dtype=float))
        p = self.full_from(x)
        v, g = self.joint(p, order=1)
        self._remember(v, p)
        return self.reduce_grad(g, self._active)

    def fn_gr(self, x=None):
        """(obj$fn(x), obj$gr(x)) from ONE evaluation -- what a quasi-Newton driver should call: its line
        search needs both at every trial point, and for the Laplace object two separate calls would run
        the inner Newton problem twice."""
        x = self.par if x is None else x
        if self._laplace is not None:
            return self._laplace.fn_gr(np.asarray(x, dtype=float))
        p = self.full_from(x)
        v, g = self.joint(p, order=1)
        self._remember(v, p)
        return v, self.reduce_grad(g, self._active)

    def _reduce_hess(self, H_full, idx):
        """Hessian w.r.t. the free groups `idx` from the full-vector Hessian (tied entries add up)."""
        ok = np.flatnonzero(self._group >= 0)
        pos = {g: k for k, g in enumerate(idx)}
        A = np.zeros((len(idx), self._full0.size))
        for i in ok:
            k = pos.get(self._group[i])
            if k is not None:
                A[k, i] = 1.0
        return A @ H_full @ A.T

    def he(self, x=None):
        """obj$he(x): exact Hessian of the objective w.r.t. obj$par (joint objects only, as in TMB;
        used by edf_conditional, R/sde.R:1363)."""
        if self._laplace is not None:
            raise NotImplementedError("Hessian is not yet implemented for models with random effects (TMB says the same)")
        x = self.par if x is None else x
        p = self.full_from(x)
        _, _, H = self.engine.hessian(p)
        self.env.last_par = p.copy()
        return self._reduce_hess(H, self._active)

    def sdreport(self, x=None, step=1e-4):
        """What SDE$fit() takes from TMB::sdreport(obj, getJointPrecision = TRUE) (R/sde.R:702-719,
        :886-887): par.fixed, par.random, cov.fixed and the joint precision of (fixed, random) in
        the order of the full free parameter vector.  Without random effects the joint precision is
        the exact Hessian.  With random effects (TMB's formula): H_f = Hessian of the Laplace
        marginal (central differences of its gradient, like optimHess), H_bb and G = H_b,theta from
        the exact joint Hessian at (theta, b_hat):  Q = [[H_f + G' H_bb^-1 G, G'], [G, H_bb]]."""
        x = np.asarray(self.par if x is None else x, dtype=float)
        if self._laplace is None:
            H = self.he(x)
            return {"par_fixed": x.copy(), "par_random": np.zeros(0), "cov_fixed": np.linalg.inv(H),
                    "jointPrecision": H, "names": list(self.names), "order": self._active.copy()}
        f, _ = self._laplace.fn_gr(x)
        p = self.env.last_par.copy()                          # (theta, b_hat)
        nt = x.size
        Hf = np.empty((nt, nt))
        for j in range(nt):
            h = step * max(1.0, abs(x[j]))
            e = np.zeros(nt)
            e[j] = h
            Hf[:, j] = (self._laplace.fn_gr(x + e)[1] - self._laplace.fn_gr(x - e)[1]) / (2 * h)
        Hf = 0.5 * (Hf + Hf.T)
        self._laplace.fn_gr(x)                                # leave b_hat / last_par at the optimum
        _, _, H = self.engine.hessian(p)
        free = np.concatenate([self._active, self._rand])
        Hj = self._reduce_hess(H, free)
        Hbb, G = Hj[nt:, nt:], Hj[nt:, :nt]
        Q = np.block([[Hf + G.T @ np.linalg.solve(Hbb, G), G.T], [G, Hbb]])
        order = np.argsort(free, kind="stable")               # full free-parameter order (fixed / random interleaved)
        return {"par_fixed": x.copy(), "par_random": p[self._rand].copy(), "cov_fixed": np.linalg.inv(Hf),
                "jointPrecision": Q[np.ix_(order, order)], "names": list(self.names_full[free][order]),
                "order": free[order], "value": f}

    def report(self, par_full=None):
        """REPORT()ed quantities at `par_full` (default: the last evaluated parameters)."""
        if self.data["type"] not in ("CTCRW", "BM_SSM", "OU_SSM"):
            return {}
        if par_full is not None:
            self.joint(np.asarray(par_full, dtype=float), order=0)
        obs = np.asarray(self.data["obs"])
        n, d = (obs.shape[0], 1) if obs.ndim == 1 else obs.shape
        return {"aest_all": self.engine.report(n, d, 2 * d if self.data["type"] == "CTCRW" else d)}

    def close(self):
        self.engine.close()


def _all_na(codes):
    codes = np.atleast_1d(np.asarray(codes, dtype=object))
    return all(c is None or (isinstance(c, float) and np.isnan(c)) for c in codes)
