"""Laplace-marginal objective for ``random = "coeff_re"`` (what TMB builds for
``MakeADFun(..., random = "coeff_re")``, R/sde.R:522-524,656-658):

    f(theta) = g(theta, b_hat) + 1/2 log det H_bb(theta, b_hat) - 1/2 n_b log(2 pi),
    b_hat    = argmin_b g(theta, b),

with g the joint penalised nllk evaluated by the CUDA engine.  The inner problem is solved by
Newton's method on the engine's analytic gradient; the Hessian block H_bb is obtained by central
differences OF THAT ANALYTIC GRADIENT (2 n_b gradient evaluations, O(h^2) error with h chosen so
that the error is ~1e-9 relative), and the outer gradient by central differences of f with
warm-started inner solves.  TMB gets both by AD-of-AD; an exact second-order adjoint for the
Kalman models is the next step of SURVEY.md 8(f) and is not built yet, so this layer is a
correct but slower stand-in: it costs O(n_b) joint evaluations per marginal evaluation.
"""
from __future__ import annotations

import numpy as np


class Laplace:
    def __init__(self, ad, newton_tol=1e-8, max_newton=50, hess_step=1e-4, grad_step=1e-5):
        self.ad = ad
        self.b = ad._full0[ad._rand].copy()
        self.newton_tol, self.max_newton = newton_tol, max_newton
        self.hess_step, self.grad_step = hess_step, grad_step
        self.n_joint_evals = 0
        self.last = None

    # ------------------------------------------------------------------------------------
    def _joint(self, x, b):
        ad = self.ad
        p = ad.full_from(x, b)
        v, g = ad.engine.eval(p, order=1)
        self.n_joint_evals += 1
        return v, ad.reduce_grad(g, ad._rand), p

    def hessian_bb(self, x, b):
        nb = b.size
        H = np.empty((nb, nb))
        for j in range(nb):
            h = self.hess_step * max(1.0, abs(b[j]))
            bp, bm = b.copy(), b.copy()
            bp[j] += h
            bm[j] -= h
            H[:, j] = (self._joint(x, bp)[1] - self._joint(x, bm)[1]) / (2 * h)
        return 0.5 * (H + H.T)

    def inner(self, x):
        """Newton iterations on b from the previous mode (warm start)."""
        b = self.b.copy()
        v, gb, p = self._joint(x, b)
        H = None
        for _ in range(self.max_newton):
            if np.max(np.abs(gb)) <= self.newton_tol * max(1.0, abs(v)):
                break
            H = self.hessian_bb(x, b)
            try:
                step = np.linalg.solve(H, gb)
            except np.linalg.LinAlgError:
                step = gb
            t = 1.0
            while True:                                   # backtracking on the joint objective
                bn = b - t * step
                vn, gbn, pn = self._joint(x, bn)
                if np.isfinite(vn) and vn <= v + 1e-12 * abs(v):
                    break
                t *= 0.5
                if t < 1e-6:
                    break
            b, v, gb, p = bn, vn, gbn, pn
        H = self.hessian_bb(x, b)                         # at the mode
        self.b = b
        return v, b, H, p

    def fn(self, x):
        v, b, H, p = self.inner(np.asarray(x, dtype=float))
        sign, logdet = np.linalg.slogdet(H)
        val = v + 0.5 * logdet - 0.5 * b.size * np.log(2 * np.pi) if sign > 0 else np.inf
        self.ad.env.last_par = p
        if np.isfinite(val) and val < self.ad.env.value_best:
            self.ad.env.value_best = val
            self.ad.env.last_par_best = p.copy()
        self.last = {"value": val, "joint": v, "b": b.copy(), "H": H}
        return val

    def gr(self, x):
        x = np.asarray(x, dtype=float)
        g = np.empty(x.size)
        b0 = self.b.copy()
        for j in range(x.size):
            h = self.grad_step * max(1.0, abs(x[j]))
            xp, xm = x.copy(), x.copy()
            xp[j] += h
            xm[j] -= h
            self.b = b0.copy()
            fp = self.fn(xp)
            self.b = b0.copy()
            fm = self.fn(xm)
            g[j] = (fp - fm) / (2 * h)
        self.b = b0
        return g
