"""Laplace marginal (value + gradient) of the OU model with a random intercept per track through the
one-pass Hessian: BASELINE configs[1] (146 random effects) and the OU half of configs[4] (8210).

    python scripts/ou_laplace.py 4096 25000
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from smoothsde_b200 import devgen
from smoothsde_b200.laplace import OnePassLaplace

T, m = int(sys.argv[1]), int(sys.argv[2])
eng, par, info = devgen.make_ou_device(T, m, device=0)
lap = OnePassLaplace(eng)
out = {"config": f"OU {T} x {m}, mu,tau ~ s(time,k=10) + s(ID,re), kappa ~ 1", "n": info["n"], "n_random": info["p_re"]}
t0 = time.perf_counter()
f, _, p = lap.eval(par, order=0)
out["laplace_value_cold_s"] = time.perf_counter() - t0
out["cold_info"] = dict(lap.info)
t0 = time.perf_counter()
f, g, p = lap.eval(p, order=1)
out["laplace_value_and_gradient_warm_s"] = time.perf_counter() - t0
out["warm_info"] = dict(lap.info)
out["marginal_nllk"] = f
outer = [i for i in range(p.size) if not (lap.o_re <= i < lap.o_re + lap.nb)]
out["gradient_outer"] = [float(g[i]) for i in outer]
# spot check of one gradient component against a central difference of the marginal value itself
k = outer[1]
h = 1e-4
vals = []
for sgn in (1.0, -1.0):
    q = p.copy(); q[k] += sgn * h
    vals.append(lap.eval(q, order=0)[0])
out["fd_check"] = {"index": k, "gradient": float(g[k]), "central_difference_of_value": (vals[0] - vals[1]) / (2 * h)}
print(json.dumps(out))
