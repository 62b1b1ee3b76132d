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


class OnePassLaplace:
    """Laplace marginal of the BM / OU models with H_bb from ONE pass over the design
    (``ssde_hess_theta_device``: X' W X + lambda S with exact per-row second-derivative blocks) instead
    of n_b tangent passes -- what makes a random intercept per track (``s(ID, bs = "re")``,
    R/sde.R:412-421; configs[1]: 146, configs[4]: 8210 random effects) tractable.  One engine handle on
    one GPU; the n_b x n_b Cholesky factorisation, solves and log-determinant run on the device
    (torch.linalg, i.e. cuSOLVER -- a plain library factorisation of the penalty block, as in
    ssde_laplace.cu).

    Value:    f(theta) = g(theta, b_hat) + 1/2 log det H_bb - n_b/2 log(2 pi), inner Newton with the exact
              H_bb (chord iterations with the previous factor first, as DeviceLaplace).
    Gradient: d f / d theta = g_theta(theta, b_hat)  [envelope theorem: g_b = 0 at the mode]
                              + d/d theta [1/2 log det H_bb(theta, b_hat(theta))],
              the second term by central differences over the non-random parameters with b_hat re-solved
              (one or two Newton steps from the mode) -- 2 x (number of non-random parameters) Hessian
              passes, independent of n_b, where the z_j-directional recipe of DeviceLaplace would need
              4 n_b tangent passes."""

    def __init__(self, engine, max_newton=50, grad_tol=1e-8, fd_step=1e-4):
        import torch
        self.engine, self._torch = engine, torch
        self.o_re, self.nb = engine.layout["coeff_re"]
        self.p_fe = engine.layout["coeff_fe"][1]
        self.p = self.p_fe + self.nb
        self.np_full = engine.n_par
        self.dev = torch.device("cuda", engine._lib.ssde_device(engine._h))
        self.max_newton, self.grad_tol, self.fd_step = max_newton, grad_tol, fd_step
        self._H = torch.zeros((self.p, self.p), dtype=torch.float64, device=self.dev)
        self._par = torch.zeros(self.np_full, dtype=torch.float64, device=self.dev)
        self._out = torch.zeros(self.np_full + 2, dtype=torch.float64, device=self.dev)
        self._L = None                     # Cholesky factor of H_bb at the previous mode (device)
        self.info = None
        # one non-default torch stream for the engine's kernels AND the torch linear algebra, so that they
        # are ordered on the device (the C ABI treats a NULL stream as the handle's own stream)
        self.stream = torch.cuda.Stream(device=self.dev)

    # ---- device helpers -------------------------------------------------------------------------------
    def _value_grad(self, p_host, order=1):
        t = self._torch
        self._par.copy_(t.as_tensor(p_host))
        try:
            self.engine.eval_device(self._par.data_ptr(), self._out.data_ptr(), order, self.stream.cuda_stream)
            out = self._out.cpu().numpy()
        except L.EngineError as e:
            if e.code != 5:
                raise
            return np.inf, None
        if out[self.np_full + 1] != 0.0:
            return np.inf, None
        return float(out[0]), (out[1:self.np_full + 1].copy() if order >= 1 else None)

    def _factor_bb(self, p_host, ridge_scale=None):
        """Cholesky factor (device) of H_bb at p_host; None if it is not positive definite.  With
        `ridge_scale` (a Newton step away from the mode): Levenberg ridge until H_bb + ridge I is."""
        t = self._torch
        self._par.copy_(t.as_tensor(p_host))
        self.engine.hess_theta_device(self._par.data_ptr(), self._H.data_ptr(), self.stream.cuda_stream)
        Hbb = self._H[self.p_fe:, self.p_fe:]
        Hs = 0.5 * (Hbb + Hbb.T)
        Lc, info = t.linalg.cholesky_ex(Hs)
        if int(info) == 0:
            return Lc, 0.0
        if ridge_scale is None:
            return None, 0.0
        # Away from the mode the exact W_i of badly fitted rows can make single coordinates concave (a
        # track's random intercept of tau, say).  First repair only those: a negative diagonal entry is
        # replaced by its absolute value; then, if needed, a global Levenberg ridge scaled to the diagonal.
        dg = t.diagonal(Hs)
        fix = t.where(dg < 0, -2.0 * dg, t.zeros_like(dg))
        scale = float(dg.abs().max())
        ridge = 0.0
        for k in range(24):
            Lc, info = t.linalg.cholesky_ex(Hs + t.diag(fix + ridge))
            if int(info) == 0:
                return Lc, max(ridge, float(fix.max()))
            ridge = 1e-8 * scale if ridge == 0.0 else ridge * 10.0
        return None, ridge

    def _solve(self, Lc, g_b):
        t = self._torch
        return t.cholesky_solve(t.as_tensor(g_b, device=self.dev)[:, None], Lc)[:, 0].cpu().numpy()

    def _mode(self, p, info):
        """Inner problem from the starting point p (full vector): returns (p at the mode, value, gradient, factor)."""
        sl = slice(self.o_re, self.o_re + self.nb)
        v, g = self._value_grad(p)
        info["n_value"] += 1
        if not np.isfinite(v):
            return p, np.inf, None, None
        tol = lambda vv: max(self.grad_tol, 1e-12 * abs(vv))
        if self._L is not None:                               # chord iterations with the previous factor
            for _ in range(12):
                gmax = float(np.max(np.abs(g[sl])))
                if gmax <= tol(v):
                    break
                q = p.copy()
                q[sl] = p[sl] - self._solve(self._L, g[sl])
                vq, gq = self._value_grad(q)
                info["n_value"] += 1
                if not np.isfinite(vq) or vq > v + 1e-14 * abs(v) or float(np.max(np.abs(gq[sl]))) > 0.5 * gmax:
                    break
                p, v, g = q, vq, gq
        Lc = None
        for it in range(self.max_newton + 1):
            gmax = float(np.max(np.abs(g[sl])))
            info["grad_max"] = gmax
            done = gmax <= tol(v) or it == self.max_newton
            # at the mode the exact H_bb must be positive definite (else the Laplace value is undefined);
            # on the way there a Levenberg ridge keeps the Newton step a descent direction
            Lc, ridge = self._factor_bb(p, None if done else gmax)
            info["n_hess"] += 1
            info["ridge_max"] = max(info.get("ridge_max", 0.0), ridge)
            if Lc is None:
                return p, np.inf, None, None
            if done:
                info["converged"] = int(gmax <= tol(v))
                break
            step = self._solve(Lc, g[sl])
            slope = float(g[sl] @ step)
            t_, ok = 1.0, False
            for _ in range(6 if gmax <= 1e3 * tol(v) else 40):
                q = p.copy()
                q[sl] = p[sl] - t_ * step
                vq, gq = self._value_grad(q)
                info["n_value"] += 1
                if np.isfinite(vq) and vq <= v - 1e-4 * t_ * slope + 1e-14 * abs(v):
                    ok = True
                    break
                t_ *= 0.5
            if not ok:
                info["converged"] = int(gmax <= 1e3 * tol(v))
                if ridge != 0.0:                              # the factor in hand is of the ridged matrix
                    Lc, _ = self._factor_bb(p, None)
                    info["n_hess"] += 1
                    if Lc is None:
                        return p, np.inf, None, None
                break
            p, v, g = q, vq, gq
            info["n_newton"] += 1
        return p, v, g, Lc

    def eval(self, par_full, order=1):
        with self._torch.cuda.stream(self.stream):
            return self._eval(par_full, order)

    def _eval(self, par_full, order):
        t = self._torch
        p0 = np.asarray(par_full, dtype=float).copy()
        info = dict(joint=np.nan, logdet=np.nan, grad_max=np.nan, converged=0, n_newton=0, n_hess=0, n_value=0, n_hvp=0)
        p, v, g, Lc = self._mode(p0, info)
        self.info = info
        if not np.isfinite(v) or Lc is None:
            return np.inf, (np.full(p0.size, np.nan) if order else None), p0
        self._L = Lc
        self._p_mode = p.copy()
        logdet = float(2.0 * t.log(t.diagonal(Lc)).sum())
        info["joint"], info["logdet"] = v, logdet
        val = v + 0.5 * logdet - 0.5 * self.nb * np.log(2 * np.pi)
        if order == 0:
            return val, None, p
        # gradient: envelope part + central differences of the profiled log-determinant
        grad = g.copy()
        sl = slice(self.o_re, self.o_re + self.nb)
        grad[sl] = 0.0
        outer = [i for i in range(p.size) if not (self.o_re <= i < self.o_re + self.nb)]
        keepL = self._L
        for k in outer:
            h = self.fd_step * max(1.0, abs(p[k]))
            ld = []
            for sgn in (+1.0, -1.0):
                q = p.copy()
                q[k] += sgn * h
                sub = dict(n_newton=0, n_hess=0, n_value=0, converged=0, grad_max=np.nan)
                self._L = keepL
                _, vq, _, Lq = self._mode(q, sub)
                info["n_hess"] += sub["n_hess"]
                info["n_value"] += sub["n_value"]
                if Lq is None or not np.isfinite(vq):
                    ld.append(np.nan)
                else:
                    ld.append(float(2.0 * t.log(t.diagonal(Lq)).sum()))
            grad[k] += 0.5 * (ld[0] - ld[1]) / (2 * h)
        self._L = keepL
        return val, grad, p

    def hessian_bb(self):
        """H_bb at the mode of the last eval (recomputed: one pass)."""
        t = self._torch
        with t.cuda.stream(self.stream):
            self._par.copy_(t.as_tensor(self._p_mode))
            self.engine.hess_theta_device(self._par.data_ptr(), self._H.data_ptr(), self.stream.cuda_stream)
            Hbb = self._H[self.p_fe:, self.p_fe:]
            return (0.5 * (Hbb + Hbb.T)).cpu().numpy()

    def close(self):
        self._H = None


class Laplace:
    """The ``fn`` / ``gr`` pair of the ``random = "coeff_re"`` object on top of an ADFun."""

    def __init__(self, ad, **opts):
        self.ad = ad
        off, size = ad.engine.layout["coeff_re"]
        if ad._rand.size != size:
            raise NotImplementedError("map on coeff_re together with random = 'coeff_re' is not supported")
        self.b = ad._full0[ad._rand].copy()
        eng = ad.engine
        drv = DeviceLaplace if hasattr(eng, "_h") else LoopLaplace
        # BM / OU on one GPU: H_bb = X' W X + lambda S from ONE pass over the design (OnePassLaplace) when the
        # library supports it for this design; tangent passes (DeviceLaplace) otherwise
        if hasattr(eng, "_h") and hasattr(eng, "supports_hess_theta") and eng.supports_hess_theta() and not opts.pop("tangent_hessian", False):
            drv = OnePassLaplace
            opts = {k: v for k, v in opts.items() if k in ("max_newton", "grad_tol", "fd_step")}
        else:
            opts.pop("tangent_hessian", None)
        self.driver = drv(eng, **opts)
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
