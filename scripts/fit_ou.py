"""SDE$fit() of the OU model with a random intercept per track (mu, tau ~ s(time, k = 10) + s(ID, bs = "re"),
kappa ~ 1): BFGS on the Laplace marginal, inner problem with the one-pass X'WX Hessian (OnePassLaplace).
BASELINE configs[1] (64 x 1e5, 146 random effects) and the OU half of configs[4] (4096 x 2.5e4, 8210).

    python scripts/fit_ou.py 4096 25000 [budget_s]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.optimize as so

from smoothsde_b200 import devgen
from smoothsde_b200.adfun import ADFun
from smoothsde_b200.laplace import OnePassLaplace

T, m = int(sys.argv[1]), int(sys.argv[2])
budget = float(sys.argv[3]) if len(sys.argv) > 3 else 300.0
eng, par, info = devgen.make_ou_device(T, m, device=0)
p_fe, n_s = info["p_fe"], info["n_s"]
par = par.copy()
par[:3] = [0.5, 0.3, 0.0]                       # truth: mu offset 0, log tau offset 0, log kappa = log 1.5
par[3:7] = 0.0                                  # log lambda (truth for the intercepts: -log 0.09 = 2.4, -log 0.04 = 3.2)
par[7:] = 0.0
pars = {"coeff_fe": par[:p_fe], "log_lambda": par[p_fe:p_fe + n_s], "coeff_re": par[p_fe + n_s:]}
obj = ADFun({"type": "OU"}, pars, random="coeff_re", engine=eng)
assert isinstance(obj._laplace.driver, OnePassLaplace)
calls = {"n": 0, "it": 0}
t0 = time.perf_counter()
trace = []


class Budget(Exception):
    pass


# The objective is handed to BFGS divided by n (what optim's control$fnscale does in R): with the identity
# as initial inverse Hessian, gradients of 1e3 .. 1e7 (nllk ~ 1.5e8) otherwise produce first steps of that size.
FSCALE = 1.0 / info["n"]


def fg(x):
    calls["n"] += 1
    if np.max(np.abs(x)) > 30.0:                  # a wild line-search trial: not worth an inner solve
        return np.inf, np.full(np.size(x), np.nan)
    f, g = obj.fn_gr(x)
    inf_ = obj._laplace.driver.info or {}
    print(f"[eval {calls['n']:3d}] {time.perf_counter() - t0:7.1f} s f = {f:.10g} |g|inf = {np.max(np.abs(g)):.4g} g = {np.array2string(np.asarray(g), precision=4)} "
          f"x = {np.array2string(np.asarray(x), precision=5)} newton {inf_.get('n_newton')} hess {inf_.get('n_hess')} conv {inf_.get('converged')} ridge {inf_.get('ridge_max', 0):.3g}",
          file=sys.stderr, flush=True)
    return f * FSCALE, g * FSCALE


def cb(xk):
    calls["it"] += 1
    el = time.perf_counter() - t0
    trace.append((el, np.asarray(xk).copy()))
    print(f"[fit ou] it {calls['it']:3d} {el:7.1f} s  evals {calls['n']}  x = {np.array2string(np.asarray(xk), precision=4)}", file=sys.stderr, flush=True)
    if el > budget:
        raise Budget()


try:
    r = so.minimize(fg, obj.par.copy(), jac=True, method="BFGS", callback=cb, options={"gtol": 1e-3 * info["n"] / 1e6 * FSCALE, "maxiter": 200})
    x, fun, jac, ok, msg = r.x, r.fun / FSCALE, r.jac / FSCALE, bool(r.success), str(r.message)
except Budget:
    x = trace[-1][1]
    fun, jac = obj.fn_gr(x)
    ok, msg = False, f"stopped by the {budget:.0f} s budget of this script"
secs = time.perf_counter() - t0
print(json.dumps({"config": f"OU {T} x {m} (n={info['n']}), mu,tau ~ s(time,k=10) + s(ID,re), kappa ~ 1; {info['p_re']} random effects",
                  "theta_names": [str(v) for v in obj.names], "theta_hat": [float(v) for v in x],
                  "kappa_hat": float(np.exp(x[2])), "sd_intercepts_hat": [float(np.exp(-0.5 * x[4])), float(np.exp(-0.5 * x[6]))],
                  "truth": {"kappa": 1.5, "sd_intercepts": [0.3, 0.2]},
                  "marginal_nllk": float(fun), "fit_wall_s": secs, "bfgs_iterations": calls["it"], "fn_gr_calls": calls["n"],
                  "success": ok, "message": msg, "grad_inf_norm": float(np.max(np.abs(jac))), "s_per_call": secs / max(calls["n"], 1),
                  "laplace_info_last": obj._laplace.driver.info}))
