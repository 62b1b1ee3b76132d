"""Laplace-marginal objective for ``random = "coeff_re"`` (what TMB builds for
``MakeADFun(..., random = "coeff_re")``, R/sde.R:522-524,656-658):

    f(theta) = g(theta, b_hat) + 1/2 log det H_bb(theta, b_hat) - 1/2 n_b log(2 pi),
    b_hat    = argmin_b g(theta, b),

with g the joint penalised nllk evaluated by the CUDA engine.

Two drivers with the same mathematics (see csrc/ssde_laplace.cu for the derivation):

* :class:`DeviceLaplace` -- one engine handle on one GPU: ``ssde_laplace_eval`` of the C ABI
  (inner Newton with the exact H_bb from tangent passes, cuSOLVER potrf / potrs on the device).
* :class:`LoopLaplace` -- any object with ``eval`` and ``hvp`` (the multi-GPU engines of
  sharded.py, whose Hessian-vector products are already all-reduced; the oracle-backed fake
  engine of the CPU tests): the same Newton / third-derivative recipe written against that
  interface, with the n_b x n_b Cholesky on the host.

The gradient of f uses  d/dx_k log det H_bb = sum_j D^3 g[z_j, z_j, e_k]  (H_bb^-1 = Z Z') with
each term a central (Richardson) difference of Hessian-vector products, and eliminates the
implicit dependence of b_hat on theta with the same factor:
``grad f = g_theta + 1/2 (w_theta - H_theta,b H_bb^-1 w_b)``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L


class DeviceLaplace:
    """ssde_laplace_* of the C ABI for one :class:`smoothsde_b200.engine.Engine`."""

    def __init__(self, engine, max_newton=100, grad_tol=1e-8, fd_step=1e-3, richardson=True):
        self.engine, self._lib = engine, engine._lib
        o = L.LaplaceOpts(int(max_newton), int(bool(richardson)), float(grad_tol), float(fd_step))
        self._w = C.c_void_p()
        rc = self._lib.ssde_laplace_create(engine._h, C.byref(o), C.byref(self._w))
        if rc != 0:
            raise L.EngineError(rc, "ssde_laplace_create failed")
        self.nb = engine.layout["coeff_re"][1]
        self.info = None

    def eval(self, par_full, order=1):
        """par_full: full vector with the starting coeff_re.  Returns (f, grad_full | None,
        par_full with coeff_re = b_hat)."""
        p = np.ascontiguousarray(par_full, dtype=np.float64).copy()
        g = np.zeros(p.size)
        val = C.c_double()
        info = L.LaplaceInfo()
        rc = self._lib.ssde_laplace_eval(self._w, p.ctypes.data_as(L.c_double_p), int(order), C.byref(val),
                                         g.ctypes.data_as(L.c_double_p), C.byref(info))
        self.info = {k: getattr(info, k) for k, _ in L.LaplaceInfo._fields_}
        if rc != 0:
            msg = self._lib.ssde_laplace_error(self._w).decode()
            if rc == 5 and "positive definite" in msg:      # not a minimum in b: f is undefined there
                return np.inf, (np.full(p.size, np.nan) if order else None), p
            raise L.EngineError(rc, msg)
        return val.value, (g if order >= 1 else None), p

    def hessian_bb(self):
        H = np.zeros((self.nb, self.nb), order="F")
        self._lib.ssde_laplace_hessian_bb(self._w, H.ctypes.data_as(L.c_double_p))
        return np.ascontiguousarray(H)

    def close(self):
        if self._w:
            self._lib.ssde_laplace_destroy(self._w)
            self._w = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class LoopLaplace:
    """The same recipe against the ``eval`` / ``hvp`` interface (multi-GPU and fake engines)."""

    def __init__(self, engine, max_newton=100, grad_tol=1e-8, fd_step=1e-3, richardson=True):
        self.engine = engine
        self.o, self.nb = engine.layout["coeff_re"]
        self.max_newton, self.grad_tol, self.fd_step, self.richardson = max_newton, grad_tol, fd_step, richardson
        self.info = None
        self._H = None
        self._Lc = None

    def _value(self, q):
        """joint value at a TRIAL point of a line search: a device-side numeric failure there (F <= 0 at
        wild parameters, SSDE_ERR_NUMERIC) means "reject the trial", as a NaN would"""
        try:
            return self.engine.eval(q, order=0)[0]
        except L.EngineError as e:
            if e.code != 5:
                raise
            return np.inf

    def _hess_cols(self, p):
        E = np.zeros((p.size, self.nb))
        E[self.o + np.arange(self.nb), np.arange(self.nb)] = 1.0
        return self.engine.hvp(p, E)                          # v, g, H[:, b]

    def eval(self, par_full, order=1):
        p = np.asarray(par_full, dtype=float).copy()
        o, nb = self.o, self.nb
        sl = slice(o, o + nb)
        info = dict(joint=np.nan, logdet=np.nan, grad_max=np.nan, converged=0, n_newton=0, n_hess=1, n_value=0, n_hvp=0)
        if nb == 0:
            v, g = self.engine.eval(p, order=order)
            info["joint"] = v
            self.info = info
            return v, g, p
        # warm start (as ssde_laplace.cu): chord iterations with the Cholesky factor of the previous mode --
        # one plain evaluation each instead of the n_b tangent passes of a Hessian
        if self._Lc is not None:
            v, g = self.engine.eval(p, order=1)
            info["n_value"] += 1
            for _ in range(12):
                gmax = float(np.max(np.abs(g[sl])))
                if not np.isfinite(gmax) or gmax <= max(self.grad_tol, 1e-12 * abs(v)):
                    break
                q = p.copy()
                q[sl] = p[sl] - np.linalg.solve(self._Lc.T, np.linalg.solve(self._Lc, g[sl]))
                vq, gq = self.engine.eval(q, order=1)
                info["n_value"] += 1
                gq_max = float(np.max(np.abs(gq[sl])))
                if not (np.isfinite(vq) and np.isfinite(gq_max)) or vq > v + 1e-14 * abs(v) or gq_max > 0.5 * gmax:
                    break
                p, v, g = q, vq, gq
        self._Lc = None
        v, g, Hc = self._hess_cols(p)
        for it in range(self.max_newton + 1):
            gb = g[sl]
            info["grad_max"] = float(np.max(np.abs(gb)))
            tol_eff = max(self.grad_tol, 1e-12 * abs(v))       # as ssde_laplace.cu: relative for objectives >= 1e4
            if info["grad_max"] <= tol_eff:
                info["converged"] = 1
                break
            if it == self.max_newton:
                break
            Hbb = 0.5 * (Hc[sl] + Hc[sl].T)
            ridge = 0.0
            while True:
                try:
                    Lc = np.linalg.cholesky(Hbb + ridge * np.eye(nb))
                    break
                except np.linalg.LinAlgError:
                    ridge = 1e-6 * (1 + info["grad_max"]) if ridge == 0.0 else 10 * ridge
            step = np.linalg.solve(Lc.T, np.linalg.solve(Lc, gb))
            slope = float(gb @ step)
            t, ok = 1.0, False
            for _ in range(6 if info["grad_max"] <= 1e3 * tol_eff else 40):
                q = p.copy()
                q[sl] = p[sl] - t * step
                vn = self._value(q)
                info["n_value"] += 1
                if np.isfinite(vn) and vn <= v - 1e-4 * t * slope + 1e-14 * abs(v):
                    ok = True
                    break
                t *= 0.5
            if not ok:
                info["converged"] = int(info["grad_max"] <= 1e3 * tol_eff)
                break
            p = q
            v, g, Hc = self._hess_cols(p)
            info["n_hess"] += 1
            info["n_newton"] += 1
        Hbb = 0.5 * (Hc[sl] + Hc[sl].T)
        self._H = Hbb
        info["joint"] = v
        try:
            Lc = np.linalg.cholesky(Hbb)
        except np.linalg.LinAlgError:
            self.info = info
            return np.inf, (np.full(p.size, np.nan) if order else None), p
        self._Lc = Lc                                          # preconditioner of the next call's chord iterations
        info["logdet"] = float(2 * np.sum(np.log(np.diag(Lc))))
        val = v + 0.5 * info["logdet"] - 0.5 * nb * np.log(2 * np.pi)
        self.info = info
        if order == 0:
            return val, None, p
        Z = np.linalg.solve(Lc.T, np.eye(nb))                  # H_bb^-1 = Z Z'
        w = np.zeros(p.size)
        eps = self.fd_step
        for j in range(nb):
            z = Z[:, j]
            nrm = np.max(np.abs(z))
            if nrm == 0:
                continue
            d = np.zeros(p.size)
            d[sl] = z / nrm

            def hv(t):
                info["n_hvp"] += 1
                return self.engine.hvp(p + t * d, d)[2]

            der = (hv(eps) - hv(-eps)) / (2 * eps)
            if self.richardson:
                der = (4 * (hv(0.5 * eps) - hv(-0.5 * eps)) / eps - der) / 3
            w += nrm * nrm * der
        u = np.linalg.solve(Lc.T, np.linalg.solve(Lc, w[sl]))
        grad = g + 0.5 * (w - Hc @ u)
        grad[sl] = 0.0
        return val, grad, p

    def hessian_bb(self):
        return self._H

    def close(self):
        pass


class Laplace:
    """The ``fn`` / ``gr`` pair of the ``random = "coeff_re"`` object on top of an ADFun."""

    def __init__(self, ad, **opts):
        self.ad = ad
        off, size = ad.engine.layout["coeff_re"]
        if ad._rand.size != size:
            raise NotImplementedError("map on coeff_re together with random = 'coeff_re' is not supported")
        self.b = ad._full0[ad._rand].copy()
        drv = DeviceLaplace if hasattr(ad.engine, "_h") else LoopLaplace
        self.driver = drv(ad.engine, **opts)
        self._cache = None                                     # (x, value, grad_active)
        self.last = None

    def _eval(self, x, order):
        x = np.asarray(x, dtype=float)
        c = self._cache
        if c is not None and np.array_equal(c[0], x) and (order == 0 or c[2] is not None):
            return c[1], c[2]
        ad = self.ad
        p0 = ad.full_from(x, self.b)                           # warm start from the previous mode
        try:
            val, g_full, p = self.driver.eval(p0, order=order)
        except L.EngineError as e:
            # a numeric failure of the engine at THIS theta (e.g. F <= 0 in the filter at wild parameters of a
            # line-search trial): the objective is undefined there, report Inf like a NaN from TMB would be
            if e.code != 5:
                raise
            val, g_full, p = np.inf, (np.full(p0.size, np.nan) if order >= 1 else None), p0
        if np.isfinite(val):
            self.b = p[ad._rand].copy()
        ad.env.last_par = p.copy()
        if np.isfinite(val) and val < ad.env.value_best:
            ad.env.value_best = val
            ad.env.last_par_best = p.copy()
        g = ad.reduce_grad(g_full, ad._active) if order >= 1 and g_full is not None else None
        self.last = dict(self.driver.info or {}, value=val, b=p[ad._rand].copy())
        self._cache = (x.copy(), val, g)
        return val, g

    def fn(self, x):
        return self._eval(x, 0)[0]

    def gr(self, x):
        return self._eval(x, 1)[1]

    def fn_gr(self, x):
        return self._eval(x, 1)

    def hessian_bb(self):
        return self.driver.hessian_bb()
