"""ctypes wrapper of the C restatement oracle (oracle/oracle_c.c).  TEST INFRASTRUCTURE / CPU
BASELINE ONLY -- see the header of oracle_c.c.  Assembles the full penalised objective and its
gradient from the C pieces (dense Kalman recursion + hand adjoint, CSR products) and numpy."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np
import scipy.sparse as sp

from . import oracle_np as O

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "liboracle.so")
_dp = C.POINTER(C.c_double)
_lib = None


def build(force=False):
    src = os.path.join(HERE, "oracle_c.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-B", "_build/liboracle.so"], stdout=subprocess.DEVNULL)
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.oracle_ctcrw.restype = C.c_double
        _lib.oracle_sde.restype = C.c_double
        _lib.oracle_max_threads.restype = C.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(_dp) if a is not None else None


class COracle:
    """Pre-converted problem so that repeated evaluations time only the objective."""

    def __init__(self, dat, nthreads=1):
        self.dat = dat
        self.type = dat["type"]
        if self.type not in ("BM", "OU", "CTCRW"):
            raise NotImplementedError("the C oracle covers BM, OU and CTCRW (use oracle_np for BM_SSM / OU_SSM)")
        self.nthreads = int(nthreads)
        obs = np.asarray(dat["obs"], dtype=float)
        self.n, self.d = obs.shape
        self.obs = np.asfortranarray(obs)
        self.ID = np.ascontiguousarray(dat["ID"], dtype=float)
        self.times = np.ascontiguousarray(dat["times"], dtype=float)
        X = sp.hstack([sp.csr_matrix(dat["X_fe"]), sp.csr_matrix(dat["X_re"])], format="csr")
        X.sum_duplicates()
        self.X = X
        self.rowptr = X.indptr.astype(np.int64)
        self.col = X.indices.astype(np.int32)
        self.val = X.data.astype(np.float64)
        self.p_fe = dat["X_fe"].shape[1]
        self.p_re = dat["X_re"].shape[1]
        if self.type == "CTCRW":
            self.a0 = np.asfortranarray(np.asarray(dat["a0"], dtype=float))
            self.P0 = np.asfortranarray(np.asarray(dat["P0"], dtype=float))
        self.n_par = X.shape[0] // self.n
        self.par_vec = np.zeros(X.shape[0])
        self.par_bar = np.zeros(X.shape[0])
        self.gth = np.zeros(X.shape[1])

    @classmethod
    def from_csr(cls, type_, ID, times, obs, rowptr, col, val, p_fe, p_re, S, ncol_re, a0=None, P0=None,
                 nthreads=1, include_penalty=1):
        """Problems too large for a scipy triplet list (the device-built 1e7-row shapes): the
        stacked design [X_fe | X_re] is given directly as CSR arrays over n_par * n rows."""
        self = cls.__new__(cls)
        self.type = type_
        self.nthreads = int(nthreads)
        self.obs = np.asfortranarray(np.asarray(obs, dtype=float))
        self.n, self.d = self.obs.shape
        self.ID = np.ascontiguousarray(ID, dtype=float)
        self.times = np.ascontiguousarray(times, dtype=float)
        self.rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
        self.col = np.ascontiguousarray(col, dtype=np.int32)
        self.val = np.ascontiguousarray(val, dtype=np.float64)
        self.p_fe, self.p_re = int(p_fe), int(p_re)
        nrow = self.rowptr.size - 1
        self.n_par = nrow // self.n

        class _Shape:                      # eval() only needs X.shape
            shape = (nrow, self.p_fe + self.p_re)
        self.X = _Shape()
        self.dat = {"type": type_, "X_fe": sp.csr_matrix((nrow, self.p_fe)), "X_re": sp.csr_matrix((nrow, self.p_re)),
                    "S": S, "ncol_re": ncol_re, "include_penalty": include_penalty}
        if type_ == "CTCRW":
            self.a0 = np.asfortranarray(np.asarray(a0, dtype=float))
            self.P0 = np.asfortranarray(np.asarray(P0, dtype=float))
        self.par_vec = np.zeros(nrow)
        self.par_bar = np.zeros(nrow)
        self.gth = np.zeros(self.p_fe + self.p_re)
        return self

    def eval(self, par, want_grad=True):
        L = lib()
        p = O.split_par(self.dat, np.asarray(par, dtype=float))
        theta = np.ascontiguousarray(np.concatenate([p["coeff_fe"], p["coeff_re"]]))
        L.oracle_spmv(C.c_int64(self.X.shape[0]), self.rowptr.ctypes.data_as(C.POINTER(C.c_int64)),
                      self.col.ctypes.data_as(C.POINTER(C.c_int32)), _p(self.val), _p(theta),
                      _p(self.par_vec), C.c_int(self.nthreads))
        gsig = C.c_double(0.0)
        pb = self.par_bar if want_grad else None
        if self.type == "CTCRW":
            v = L.oracle_ctcrw(C.c_int64(self.n), C.c_int(self.d), _p(self.ID), _p(self.times),
                               _p(self.obs), _p(self.par_vec), _p(self.a0),
                               C.c_int64(self.a0.shape[0]), _p(self.P0),
                               C.c_double(float(p["log_sigma_obs"])), _p(pb), C.byref(gsig), None,
                               C.c_int(self.nthreads))
            pen = O.penalty_kalman(self.dat, p["log_lambda"], p["coeff_re"])
        else:
            v = L.oracle_sde(C.c_int(0 if self.type == "BM" else 1), C.c_int64(self.n), C.c_int(self.d),
                             _p(self.ID), _p(self.times), _p(self.obs), _p(self.par_vec), _p(pb),
                             C.c_int(self.nthreads))
            pen = O.penalty_sde(self.dat, p["log_lambda"], p["coeff_re"])
        nllk = v + pen
        if not want_grad:
            return nllk, None
        L.oracle_spmv_t(C.c_int64(self.X.shape[0]), C.c_int64(self.X.shape[1]),
                        self.rowptr.ctypes.data_as(C.POINTER(C.c_int64)),
                        self.col.ctypes.data_as(C.POINTER(C.c_int32)), _p(self.val), _p(self.par_bar),
                        _p(self.gth))
        g_fe = self.gth[:self.p_fe].copy()
        g_re = self.gth[self.p_fe:].copy()
        ncol_re = np.atleast_1d(np.asarray(self.dat["ncol_re"], dtype=np.int64))
        g_ll = np.zeros(p["log_lambda"].size)
        use_pen = ncol_re[0] > 0 and (self.type == "CTCRW" or int(self.dat.get("include_penalty", 1)) != 0)
        if use_pen:
            S = self.dat["S"]
            o = 0
            for i, Sn in enumerate(ncol_re):
                Sn = int(Sn)
                b = p["coeff_re"][o:o + Sn]
                Sb = np.asarray(O._block(S, o, o + Sn) @ b).ravel()
                lam = np.exp(p["log_lambda"][i])
                g_re[o:o + Sn] += lam * Sb
                g_ll[i] = -0.5 * Sn + 0.5 * lam * (b @ Sb)
                o += Sn
        pieces = ([np.array([gsig.value])] if self.type == "CTCRW" else []) + [g_fe, g_ll, g_re]
        return nllk, np.concatenate(pieces)

    def hessian(self, par, k=1e-3):
        """Joint Hessian of the penalised objective: Richardson-extrapolated central differences
        (error O(k^4)) of the analytic gradient above.  TMB gets it by AD-of-AD."""
        par = np.asarray(par, dtype=float)
        n = par.size
        H = np.empty((n, n))
        for j in range(n):
            e = np.zeros(n)
            e[j] = 1.0
            d1 = (self.eval(par + k * e)[1] - self.eval(par - k * e)[1]) / (2 * k)
            d2 = (self.eval(par + 0.5 * k * e)[1] - self.eval(par - 0.5 * k * e)[1]) / k
            H[:, j] = (4 * d2 - d1) / 3
        return 0.5 * (H + H.T)

    def aest(self, par):
        L = lib()
        p = O.split_par(self.dat, np.asarray(par, dtype=float))
        theta = np.ascontiguousarray(np.concatenate([p["coeff_fe"], p["coeff_re"]]))
        L.oracle_spmv(C.c_int64(self.X.shape[0]), self.rowptr.ctypes.data_as(C.POINTER(C.c_int64)),
                      self.col.ctypes.data_as(C.POINTER(C.c_int32)), _p(self.val), _p(theta),
                      _p(self.par_vec), C.c_int(self.nthreads))
        out = np.zeros((self.n, 2 * self.d), order="F")
        L.oracle_ctcrw(C.c_int64(self.n), C.c_int(self.d), _p(self.ID), _p(self.times), _p(self.obs),
                       _p(self.par_vec), _p(self.a0), C.c_int64(self.a0.shape[0]), _p(self.P0),
                       C.c_double(float(p["log_sigma_obs"])), None, None, _p(out), C.c_int(1))
        return np.ascontiguousarray(out)


def max_threads():
    return lib().oracle_max_threads()
