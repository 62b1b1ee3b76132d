"""Python mirror of the R6 class ``SDE`` (R/sde.R:16-1818) for the part of its interface that
sits on the hot path: ``SDE(formulas, data, type, response, par0, fixpar, other_data)``,
``setup()``, ``fit()``, ``logLik()``, coefficient accessors / mutators.  R is not available in
this image, so this module plays the role of ``R/sde.R`` above the C ABI: it builds exactly the
data list / parameter list / ``map`` that ``SDE$setup()`` hands to ``TMB::MakeADFun``
(R/sde.R:491-670) and drives the outer optimisation like ``SDE$fit()`` (R/sde.R:683-720).
Same argument meaning and error behaviour as the reference; presentation, prediction and
posterior code are out of scope.
"""
from __future__ import annotations

import warnings
from collections import OrderedDict

import numpy as np
import scipy.sparse as sp

from . import design as _design
from .adfun import ADFun

# link / inverse link per SDE type (R/sde.R:56-87); only the built types are listed
_LINKS = {
    "BM": lambda d: [("mu", None)] * d + [("sigma", "log")],
    "OU": lambda d: [("mu", None)] * d + [("tau", "log"), ("kappa", "log")],
    "CTCRW": lambda d: [("mu", None)] * d + [("tau", "log"), ("nu", "log")],
    "BM_SSM": lambda d: [("mu", None)] * d + [("sigma", "log")],                       # R/sde.R:60-63
    "OU_SSM": lambda d: [("mu", None)] * d + [("tau", "log"), ("kappa", "log")],       # R/sde.R:68-71
}
_KNOWN_UNBUILT = ("BM_t", "CIR", "ESEAL_SSM")


def _par_names(type_, n_dim):
    spec = _LINKS[type_](n_dim)
    out = []
    k = 0
    for nm, link in spec:
        if nm == "mu":
            k += 1
            out.append((f"mu{k}" if n_dim > 1 else "mu", link))
        else:
            out.append((nm, link))
    return out


class SDE:
    def __init__(self, formulas=None, data=None, type=None, response=None, par0=None, fixpar=None,
                 other_data=None, device=0, adfun_factory=None):
        if isinstance(response, str):
            response = [response]
        self._type, self._response, self._fixpar = type, list(response), list(fixpar or [])
        data = {k: np.asarray(v) for k, v in dict(data).items()}
        if any(r not in data for r in self._response):
            raise ValueError("'response' not found in 'data'")                  # R/sde.R:51-52
        if type not in _LINKS:
            if type in _KNOWN_UNBUILT:
                raise NotImplementedError(f"SDE type '{type}' exists in smoothSDE but is not built here")
            raise ValueError("Unknown SDE type")                                # src/smoothSDE.cpp:25
        n_dim = len(self._response)
        spec = _par_names(type, n_dim)
        names = [nm for nm, _ in spec]
        self._links = OrderedDict(spec)
        if formulas is None:
            formulas = OrderedDict((nm, "~ 1") for nm in names)
        else:
            formulas = OrderedDict(formulas)
            if len(formulas) != len(names):                                     # R/sde.R:93-98
                raise ValueError(f"'formulas' should be a list of length {len(names)} for the model {type}, "
                                 f"with components {', '.join(names)}")
            if list(formulas.keys()) != names:                                  # R/sde.R:99-102
                raise ValueError(f"'formulas' should be a list with components {', '.join(names)}")
        for nm in self._fixpar:                                                 # R/sde.R:103-105
            if formulas[nm].replace(" ", "") != "~1":
                raise ValueError("formulas should be ~1 for fixed parameters")
        self._formulas = formulas
        n = len(next(iter(data.values())))
        if "ID" not in data:                                                    # R/sde.R:108-114
            warnings.warn("No ID column found in 'data', assuming same ID for all observations")
            data["ID"] = np.ones(n, dtype=np.int64)
        if "time" not in data:                                                  # R/sde.R:117-119
            raise ValueError("'data' should have a time column")
        self._data = data
        self._other_data = dict(other_data or {})
        self._device = device
        self._adfun_factory = adfun_factory or ADFun
        self._mats = _design.make_design(formulas, data, n)                     # make_mat, R/sde.R:378-455
        m = self._mats
        self._coeff_fe = np.zeros(m.X_fe.shape[1])                              # R/sde.R:138-140
        self._coeff_re = np.zeros(m.X_re.shape[1])
        self._lambda = np.ones(len(m.ncol_re))
        if par0 is not None:                                                    # R/sde.R:143-160
            par0 = np.atleast_1d(np.asarray(par0, dtype=float))
            if par0.size != len(names):
                raise ValueError(f"'par0' should be of length {len(names)} with one entry for each SDE parameter "
                                 f"({', '.join(names)})")
            i0 = np.r_[0, np.cumsum(m.ncol_fe)[:-1]]
            for k, (nm, link) in enumerate(spec):
                self._coeff_fe[i0[k]] = np.log(par0[k]) if link == "log" else par0[k]
        self._tmb_obj = self._tmb_obj_joint = None
        self._out = None
        # ---- decay terms, R/sde.R:162-180
        od = self._other_data
        if od.get("t_decay") is not None:
            if od.get("col_decay") is None:                                     # R/sde.R:165-169
                term = od["decay_term"]
                od["col_decay"] = np.array([i + 1 for i, nm in enumerate(m.names_re) if nm[:len(term)] == term])
            if np.size(od["t_decay"]) != len(formulas) * n:                     # R/sde.R:170-173
                raise ValueError("'other_data$t_decay' should be of length (number of parameters) x (number of data)")
            if np.size(od["col_decay"]) != np.size(od["ind_decay"]):            # R/sde.R:174-176
                raise ValueError("Check length of 'other_data$ind_decay' and 'other_data$col_decay'")
            self._rho = np.ones(np.unique(np.asarray(od["ind_decay"])).size)    # R/sde.R:177
        else:
            self._rho = np.ones(1)

    def rho(self): return self._rho
    def update_rho(self, v): self._rho = np.atleast_1d(np.asarray(v, dtype=float)).copy()      # R/sde.R:358-360
    def other_data(self): return self._other_data

    def X_re_decay(self):
        """X_re with the decaying columns multiplied by exp(-rho * t_decay), R/sde.R:303-326."""
        od = self._other_data
        if od.get("t_decay") is None:
            raise ValueError("This model has no decaying terms")               # R/sde.R:322
        X = sp.lil_matrix(sp.csc_matrix(self._mats.X_re, dtype=float))
        Xc = sp.csc_matrix(self._mats.X_re)
        t = np.asarray(od["t_decay"], dtype=float)
        for col, ind in zip(np.atleast_1d(od["col_decay"]), np.atleast_1d(od["ind_decay"])):
            c = Xc[:, int(col) - 1]
            X[c.indices, int(col) - 1] = (c.data * np.exp(-self._rho[int(ind) - 1] * t[c.indices])).reshape(-1, 1)
        return sp.csr_matrix(X)

    # ---- accessors / mutators (R/sde.R:186-360)
    def formulas(self): return self._formulas
    def data(self): return self._data
    def type(self): return self._type
    def response(self): return self._response
    def fixpar(self): return self._fixpar
    def mats(self): return self._mats
    def coeff_fe(self): return self._coeff_fe
    def coeff_re(self): return self._coeff_re
    def lambda_(self): return self._lambda
    def sdev(self): return 1.0 / np.sqrt(self._lambda)
    def out(self): return self._out
    def tmb_obj(self): return self._tmb_obj
    def tmb_obj_joint(self): return self._tmb_obj_joint

    def update_coeff_fe(self, v): self._coeff_fe = np.asarray(v, dtype=float).copy()
    def update_coeff_re(self, v): self._coeff_re = np.asarray(v, dtype=float).copy()
    def update_lambda(self, v): self._lambda = np.asarray(v, dtype=float).copy()

    def obs(self):
        return np.column_stack([np.asarray(self._data[r], dtype=float) for r in self._response])

    def ind_fixcoeff(self):
        """Indices in coeff_fe of the (intercept) coefficients of fixed parameters."""
        i0 = np.r_[0, np.cumsum(self._mats.ncol_fe)[:-1]]
        names = list(self._formulas.keys())
        return [int(i0[names.index(nm)]) for nm in self._fixpar]

    # ---- the two MakeADFun calls of SDE$setup(), R/sde.R:491-670
    def tmb_lists(self, map=None):
        m = self._mats
        n = len(self._data["time"])
        map = dict(map or {})
        has_re = m.S is not None and m.X_re.shape[1] > 0
        tmb_par = OrderedDict(coeff_fe=self._coeff_fe.copy(), log_lambda=np.zeros(1), log_decay=np.log(self._rho),
                              coeff_re=np.zeros(1))                             # R/sde.R:504-507
        random = None
        if not has_re:                                                          # R/sde.R:511-518
            map["coeff_re"] = [None]
            map["log_lambda"] = [None]
            S = sp.csr_matrix((1, 1))
            ncol_re = np.array([0])
            X_re = sp.csr_matrix((m.X_fe.shape[0], 1))
        else:                                                                   # R/sde.R:519-525
            random = "coeff_re"
            tmb_par["coeff_re"] = self._coeff_re.copy()
            tmb_par["log_lambda"] = np.log(self._lambda)
            S, ncol_re, X_re = m.S, np.asarray(m.ncol_re), m.X_re
        _, first = np.unique(self._data["ID"], return_index=True)
        codes, _ = _design.factor_codes(self._data["ID"])
        tmb_dat = {"type": self._type, "ID": (codes + 1).astype(float), "times": np.asarray(self._data["time"], float),
                   "obs": self.obs(), "X_fe": m.X_fe, "X_re": X_re, "S": S, "ncol_re": ncol_re,
                   "include_penalty": 1}
        if self._type in ("BM_SSM", "OU_SSM"):                                  # R/sde.R:542-568
            ID = tmb_dat["ID"]
            i0 = np.r_[0, np.nonzero(ID[:-1] != ID[1:])[0] + 1]
            tmb_dat["a0"] = tmb_dat["obs"][i0].copy()                           # first observation of each track
            P0 = self._other_data.get("P0")
            tmb_dat["P0"] = 10.0 * np.eye(len(self._response)) if P0 is None else np.asarray(P0, float)
            tmb_par = OrderedDict([("log_sigma_obs", np.zeros(1))] + list(tmb_par.items()))
            if self._other_data.get("H") is not None:
                tmb_dat["H_array"] = np.asarray(self._other_data["H"], float)
                map["log_sigma_obs"] = [None]
        if self._type == "CTCRW":                                               # R/sde.R:569-598
            n_dim = len(self._response)
            ID = tmb_dat["ID"]
            i0 = np.r_[0, np.nonzero(ID[:-1] != ID[1:])[0] + 1]
            a0 = np.zeros((i0.size, 2 * n_dim))
            for i in range(n_dim):
                a0[:, 2 * i] = tmb_dat["obs"][i0, i]
            tmb_dat["a0"] = a0
            P0 = self._other_data.get("P0")
            tmb_dat["P0"] = np.diag(np.tile([1.0, 10.0], n_dim)) if P0 is None else np.asarray(P0, float)
            tmb_par = OrderedDict([("log_sigma_obs", np.zeros(1))] + list(tmb_par.items()))
            if self._other_data.get("H") is not None:
                tmb_dat["H_array"] = np.asarray(self._other_data["H"], float)
                map["log_sigma_obs"] = [None]
        if self._type in ("BM", "OU"):                                          # R/sde.R:635-649
            od = self._other_data
            if od.get("t_decay") is not None:
                if np.any(np.asarray(od["col_decay"]) > len(m.names_re)):
                    raise ValueError(f"'col_decay' should be between 1 and {len(m.names_re)}")
                tmb_dat["t_decay"] = np.asarray(od["t_decay"], float)
                tmb_dat["col_decay"] = np.asarray(od["col_decay"], np.int64)
                tmb_dat["ind_decay"] = np.asarray(od["ind_decay"], np.int64)
            else:
                tmb_dat["t_decay"], tmb_dat["col_decay"], tmb_dat["ind_decay"] = np.zeros(1), np.zeros(1, int), np.zeros(1, int)
                map["log_decay"] = [None]
        else:
            tmb_par.pop("log_decay")                                            # R/sde.R:650-653
        if self._fixpar:                                                        # R/sde.R:621-632
            cmap = list(range(m.X_fe.shape[1]))
            for i in self.ind_fixcoeff():
                cmap[i] = None
            map["coeff_fe"] = cmap
        return tmb_dat, tmb_par, map, random

    def setup(self, silent=True, map=None):
        tmb_dat, tmb_par, map, random = self.tmb_lists(map)
        self._tmb_obj = self._adfun_factory(tmb_dat, tmb_par, map=map, random=random, device=self._device)
        dat_joint = dict(tmb_dat, include_penalty=0)                            # R/sde.R:665-669
        self._tmb_obj_joint = self._adfun_factory(dat_joint, tmb_par, map=map, random=None, device=self._device)

    # ---- SDE$fit(), R/sde.R:683-720: optim(par, fn, gr, method = "BFGS")
    def fit(self, silent=True, map=None, gtol=1e-5, maxiter=200, sdreport=False):
        from scipy.optimize import minimize
        if self._tmb_obj is None:
            self.setup(silent=silent, map=map)
        obj = self._tmb_obj
        # optim(fn, gr, method = "BFGS") in R/sde.R:694-697; value and gradient come from one evaluation
        res = minimize(obj.fn_gr, obj.par, jac=True, method="BFGS", options={"gtol": gtol, "maxiter": maxiter})
        self._out = res
        p = obj.env.last_par_best
        lay = obj.engine.layout
        off, size = lay["coeff_fe"]
        self.update_coeff_fe(p[off:off + size])
        if len(self._mats.ncol_re) > 0:
            off, size = lay["coeff_re"]
            self.update_coeff_re(p[off:off + size])
            off, size = lay["log_lambda"]
            self.update_lambda(np.exp(p[off:off + size]))
        if "log_decay" in lay:                                                  # R/sde.R:715-719
            off, size = lay["log_decay"]
            self.update_rho(np.exp(p[off:off + size]))
        self._par_all = p.copy()
        if sdreport:                                   # R/sde.R:702-704: sdreport(obj, getJointPrecision = TRUE)
            self._rep = obj.sdreport(res.x)
        return res

    # ---- logLik.SDE, R/utility.R:115-123: -tmb_obj_joint$fn(par_all)
    def logLik(self):
        if self._tmb_obj_joint is None or self._out is None:
            raise RuntimeError("fit the model first")
        obj = self._tmb_obj_joint
        free = obj._free
        return -obj.fn(self._par_all[obj._active]) if obj._rand.size == 0 else -obj.joint(self._par_all, 0)[0]
