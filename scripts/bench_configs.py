"""Timings of the BASELINE.json configurations other than the headline one, plus second-order
work, on one GPU.  Writes one JSON object per line (profiles/r01_configs.jsonl when redirected).

    python scripts/bench_configs.py [--skip-ou]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from smoothsde_b200 import devgen, synth
from smoothsde_b200 import sharded as S
from smoothsde_b200.engine import Engine
from smoothsde_b200.laplace import DeviceLaplace

PEAK = 6650.0
pp = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(pp):
    PEAK = float(json.load(open(pp))["hbm_gbs"])


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    return float(np.median(ts))


def report(name, n, ms, b_alg, extra=None):
    line = {"config": name, "n": n, "ms_per_eval": ms, "obs_eval_per_s": n / ms * 1e3, "alg_bytes_per_obs": b_alg,
            "achieved_GBs": b_alg * n / ms / 1e6, "frac_of_hbm_peak": b_alg * n / ms / 1e6 / PEAK, "peak_GBs": PEAK}
    line.update(extra or {})
    print(json.dumps(line), flush=True)


def ctcrw_case(name, T, m, **kw):
    eng, par, info = devgen.make_ctcrw_device(T, m, device=0, **kw)
    n = info["n"]
    b_alg = devgen.alg_bytes_per_obs(2, 4, 22)
    ms = timed(lambda: eng.eval(par, 1))
    d = np.zeros(par.size)
    d[-1] = 1.0
    ms_h = timed(lambda: eng.hvp(par, d), reps=5, warm=2)
    extra = {"kernel_ms_last_eval": None, "hvp_ms_per_direction": ms_h}
    lap = DeviceLaplace(eng)
    t0 = time.perf_counter()
    f, g, p = lap.eval(par, order=1)
    extra["laplace_value_and_gradient_s"] = time.perf_counter() - t0
    extra["laplace_info"] = lap.info
    t0 = time.perf_counter()
    f2, g2, p2 = lap.eval(p, order=1)                    # warm start at the mode: what every later BFGS step costs
    extra["laplace_warm_s"] = time.perf_counter() - t0
    extra["laplace_warm_info"] = lap.info
    lap.close()
    report(name, n, ms, b_alg, extra)
    return eng, par, info


def main():
    # configs[2]: the headline shape, for reference next to the others
    eng, par, info = ctcrw_case("C3 CTCRW 1024 x 1e5 (headline)", 1024, 100000)
    eng.close(); torch.cuda.empty_cache()
    # configs[4] (CTCRW half): many short tracks
    eng, par, info = ctcrw_case("C5 CTCRW 4096 x 2.5e4", 4096, 25000)
    eng.close(); torch.cuda.empty_cache()
    # configs[3]: one track of 1e8 rows; single handle and 8 time shards on the same arrays
    n = 1024 * 97656
    eng, par, info = ctcrw_case("C4 CTCRW 1 x 1e8 (single handle)", 1, n, sim_tracks=1024)
    per = (n // 8) // 1024 * 1024
    slabs = devgen.slab_views(info, eng, [r * per for r in range(8)] + [n])
    ts = S.TimeShardedEngine.from_engines(slabs, [0] * 8)
    ms = timed(lambda: ts.eval(par), reps=5, warm=2)
    report("C4 CTCRW 1 x 1e8, 8 time shards driven from one process on ONE GPU (protocol overhead)", n, ms,
           devgen.alg_bytes_per_obs(2, 4, 22))
    for e in slabs:
        e.close()
    eng.close(); torch.cuda.empty_cache()
    if "--skip-ou" in sys.argv:
        return
    # configs[1] and the OU half of configs[4]: OU, mu, tau ~ s(time) + s(ID, bs = "re"), kappa ~ 1, built on the device
    for name, T, m, with_laplace in (("C2 OU 64 x 1e5, s(time) + s(ID, re)  [configs[1]]", 64, 100000, True),
                                     ("C5 OU 4096 x 2.5e4, s(time) + s(ID, re)  [configs[4], OU half]", 4096, 25000, False)):
        eng, par, info = devgen.make_ou_device(T, m, device=0)
        n = info["n"]
        b_alg = devgen.alg_bytes_per_obs(1, 3, 23)
        ms = timed(lambda: eng.eval(par, 1))
        d = np.zeros(par.size)
        d[-1] = 1.0
        ms_h = timed(lambda: eng.hvp(par, d), reps=5, warm=2)
        sb = info["stored_bytes_per_obs"]
        extra = {"p_re": info["p_re"], "stored_bytes_per_obs": sb, "kernel_ms_last_eval": eng.last_eval_ms,
                 "dram_frac_stored_bytes": sb * n / (eng.last_eval_ms * 1e-3) / 1e9 / PEAK, "hvp_ms_per_direction": ms_h}
        if with_laplace:                     # 146 random effects: the dense Laplace driver (146 tangent passes per Hessian)
            lap = DeviceLaplace(eng)
            t0 = time.perf_counter()
            f, g, p = lap.eval(par, order=1)
            extra["laplace_value_and_gradient_s"] = time.perf_counter() - t0
            extra["laplace_info"] = lap.info
            t0 = time.perf_counter()
            lap.eval(p, order=1)
            extra["laplace_warm_s"] = time.perf_counter() - t0
            extra["laplace_warm_info"] = lap.info
            lap.close()
        else:
            extra["laplace"] = "not built at this size: H_bb has arrowhead structure (18 spline coefficients + 2 x 4096 intercepts), DESIGN.md section 8"
        report(name, n, ms, b_alg, extra)
        eng.close(); torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
