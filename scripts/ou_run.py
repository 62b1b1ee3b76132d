"""OU with random intercepts per track on the device (BASELINE configs[1] / OU half of configs[4]):
a few nllk + gradient evaluations, for timing and ncu captures.

    python scripts/ou_run.py 4096 25000 [reps]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from smoothsde_b200 import devgen

T, m = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
eng, par, info = devgen.make_ou_device(T, m, device=0)
eng.set_profile(True)
ms = []
for _ in range(reps):
    v, g = eng.eval(par, 1)
    ms.append(dict(eng.last_kernel_times()))
n = info["n"]
k = {nm: float(np.median([d[nm] for d in ms[1:] or ms])) for nm in ms[0]}
tot = sum(k.values())
print(json.dumps({"config": f"OU {T} x {m}, mu,tau ~ s(time,k=10) + s(ID,re), kappa ~ 1", "n": n, "p_re": info["p_re"],
                  "nllk": v, "kernels_ms": k, "ms_per_eval": tot, "obs_eval_per_s": n / tot * 1e3,
                  "alg_bytes_per_obs": devgen.alg_bytes_per_obs(1, 3, 23),
                  "stored_bytes_per_obs": info["stored_bytes_per_obs"]}))
eng.close()
