"""Tiny standalone run of every kernel family, used under compute-sanitizer (not collected by pytest):

    compute-sanitizer --tool memcheck  python tests/gpu_smoke_small.py
    compute-sanitizer --tool racecheck python tests/gpu_smoke_small.py

Covers: ctcrw_fwd / ctcrw_bwd (decoupled, several tiles so that the look-back runs), sde_stream (OU with
random intercepts, BM), sde_fused (tangent pass), sde_decay, the coupled filter (user H_array), the
tangent kernels (hvp), time shards on one device, the Laplace driver."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np

from oracle import oracle_np as O
from smoothsde_b200 import sharded, synth
from smoothsde_b200.engine import Engine
from smoothsde_b200.laplace import DeviceLaplace

for model, T, m, nd, re_id in [("CTCRW", 2, 2700, 2, None), ("OU", 3, 300, 1, True), ("BM", 1, 150, 2, None), ("OU_SSM", 2, 300, 1, None)]:
    dat, par, info = synth.make_problem(model, T, m, missing_frac=0.1, n_dim=nd, re_id=re_id)
    eng = Engine.from_data(dat)
    v, g = eng.eval(par, 1)
    ref = O.nllk(dat, par)
    d = np.zeros(par.size); d[-1] = 1.0
    _, _, hv = eng.hvp(par, d)
    print(model, v, ref, abs(v - ref) / abs(ref), "ms", eng.last_eval_ms, "launches", eng.last_eval_launches, "hv", float(np.abs(hv).max()))
    eng.close()
# decay terms
from test_decay import decay_problem
dat, par, info = decay_problem("OU", 2, 120, 1, 5)
eng = Engine.from_data(dat)
print("decay", eng.eval(par, 1)[0], O.nllk(dat, par))
eng.close()
# coupled filter
dat, par, info = synth.make_problem("CTCRW", 2, 300, missing_frac=0.05, n_dim=2, seed=12, k=5)
rng = np.random.default_rng(0)
A = rng.normal(size=(info["n"], 2, 2)) * 0.1
dat["H_array"] = np.ascontiguousarray((A @ A.transpose(0, 2, 1) + 0.01 * np.eye(2)).transpose(1, 2, 0))
eng = Engine.from_data(dat)
print("coupled", eng.eval(par, 1)[0])
eng.close()
# time shards on one device + Laplace
dat, par, info = synth.make_problem("CTCRW", 1, 3000, missing_frac=0.05, n_dim=2, seed=9)
ts = sharded.TimeShardedEngine(dat, devices=[0, 0, 0])
print("time shards", ts.eval(par)[0], O.nllk(dat, par))
ts.close()
dat, par, info = synth.make_problem("CTCRW", 2, 150, n_dim=2, seed=40, k=5)
eng = Engine.from_data(dat)
lap = DeviceLaplace(eng)
f, g, p = lap.eval(par, order=1)
print("laplace", f, lap.info["converged"])
lap.close(); eng.close()
