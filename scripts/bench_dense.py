"""Throughput of the coupled filter (user H_array) and of a decay model on one GPU, host-built
problems of moderate size.  One JSON object per line.

    python scripts/bench_dense.py [tracks] [steps]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch

from smoothsde_b200 import synth
from smoothsde_b200.engine import Engine


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append((time.perf_counter() - t0) * 1e3)
    return float(np.median(ts))


def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    m = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
    rng = np.random.default_rng(1)
    cases = () if "--only-decay" in sys.argv else (("CTCRW", 2), ("OU_SSM", 2), ("CTCRW", 1))
    for model, nd in cases:
        dat, par, info = synth.make_problem(model, T, m, missing_frac=0.05, n_dim=nd, seed=3)
        n = info["n"]
        e0 = Engine.from_data(dat)
        ms0 = timed(lambda: e0.eval(par, 1))
        e0.close()
        A = rng.normal(size=(n, nd, nd)) * 0.05
        dat["H_array"] = np.ascontiguousarray((A @ A.transpose(0, 2, 1) + 0.01 * np.eye(nd)).transpose(1, 2, 0))
        e1 = Engine.from_data(dat)
        ms1 = timed(lambda: e1.eval(par, 1))
        d = np.zeros(par.size); d[-1] = 1.0
        msh = timed(lambda: e1.hvp(par, d), reps=3, warm=1)
        e1.close()
        print(json.dumps({"config": f"{model} d={nd}, {T} x {m} rows", "n": n, "decoupled_ms": ms0, "coupled_user_H_ms": ms1,
                          "coupled_obs_eval_per_s": n / ms1 * 1e3, "coupled_hvp_ms": msh}), flush=True)
    from test_decay import decay_problem
    dat, full, info = decay_problem("OU", T, m, 1, 5, 2)
    n = dat["obs"].shape[0]
    plain = {k: v for k, v in dat.items() if k not in ("t_decay", "col_decay", "ind_decay")}
    o = info["p_fe"] + np.atleast_1d(dat["ncol_re"]).size
    e0 = Engine.from_data(plain)
    ms0 = timed(lambda: e0.eval(np.r_[full[:o], full[o + 2:]], 1))
    e0.close()
    e1 = Engine.from_data(dat)
    ms1 = timed(lambda: e1.eval(full, 1))
    e1.close()
    print(json.dumps({"config": f"OU d=1 with decay terms, {T} x {m} rows", "n": n, "fused_ms": ms0, "decay_ms": ms1,
                      "decay_obs_eval_per_s": n / ms1 * 1e3}), flush=True)


if __name__ == "__main__":
    main()
