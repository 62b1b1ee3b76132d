"""One small run of every kernel family, for ncu captures.

    ncu --set full --clock-control none --import-source on -k regex:'sde_fused|sde_decay|ctcrw_fwd|ctcrw_bwd' \
        -o gpurun_out/families python scripts/kernel_families.py

Families: sde_fused_kernel (OU, s(time) + s(ID, re)), sde_decay_kernel (OU with decay terms),
ctcrw_fwd/bwd on DenseModel (CTCRW with a user H_array: coupled 4-state filter), ctcrw_fwd/bwd on
Dual (Hessian-vector product of the decoupled CTCRW model).  Prints, per family, the algorithmic
bytes per row (SURVEY 8(d) accounting) so the summaries can quote `frac` beside the ncu counters."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np

from smoothsde_b200 import devgen, synth
from smoothsde_b200.engine import Engine


def alg(dat, n, nd, n_par):
    nnz = (dat["X_fe"].nnz + dat["X_re"].nnz) / n
    return {"nnz_per_row": nnz, "alg_bytes_per_obs": devgen.alg_bytes_per_obs(nd, n_par, nnz)}


def main():
    out = {}
    # BM / OU fused map-reduce: config[1] shape at 64 x 2e4 rows
    dat, par, info = synth.make_problem("OU", 64, 20000, n_dim=1, seed=20260102, re_id=True)
    eng = Engine.from_data(dat)
    for _ in range(3):
        eng.eval(par, 1)
    out["sde_fused (OU 64 x 2e4, s(time)+s(ID,re))"] = dict(alg(dat, info["n"], 1, 3), n=info["n"], kernel_ms=eng.last_eval_ms)
    d = np.zeros(par.size); d[-1] = 1.0
    eng.hvp(par, d)
    eng.close()
    # decay terms
    from test_decay import decay_problem
    dat, par = decay_problem("OU", 64, 20000, 1, 113)[:2]
    eng = Engine.from_data(dat)
    for _ in range(3):
        eng.eval(par, 1)
    n = dat["obs"].shape[0]
    out["sde_decay (OU 64 x 2e4 with decay terms)"] = dict(alg(dat, n, 1, 3), n=n, kernel_ms=eng.last_eval_ms)
    eng.close()
    # coupled filter: CTCRW d = 2 with per-row measurement covariances
    dat, par, info = synth.make_problem("CTCRW", 64, 20000, n_dim=2, seed=12)
    rng = np.random.default_rng(0)
    A = rng.normal(size=(info["n"], 2, 2)) * 0.1
    dat["H_array"] = np.ascontiguousarray((A @ A.transpose(0, 2, 1) + 0.01 * np.eye(2)).transpose(1, 2, 0))
    eng = Engine.from_data(dat)
    for _ in range(3):
        eng.eval(par, 1)
    a = alg(dat, info["n"], 2, 4)
    a["alg_bytes_per_obs"] += 2 * 3 * 8           # H_i: 3 doubles per row, read by both kernels
    out["DenseModel fwd/bwd (CTCRW 64 x 2e4, user H_array)"] = dict(a, n=info["n"], kernel_ms=eng.last_eval_ms)
    eng.close()
    # tangent (Dual) pair on the decoupled model
    dat, par, info = synth.make_problem("CTCRW", 64, 20000, n_dim=2, seed=13)
    eng = Engine.from_data(dat)
    d = np.zeros(par.size); d[-1] = 1.0
    for _ in range(3):
        eng.hvp(par, d)
    out["Dual fwd/bwd (CTCRW 64 x 2e4, one Hessian-vector product)"] = dict(alg(dat, info["n"], 2, 4), n=info["n"])
    eng.eval(par, 1)
    eng.close()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
