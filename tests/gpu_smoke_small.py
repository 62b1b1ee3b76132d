"""Tiny standalone check used under compute-sanitizer (not collected by pytest)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import oracle_np as O
from smoothsde_b200 import synth
from smoothsde_b200.engine import Engine
for model, T, m, nd in [("CTCRW", 2, 700, 2), ("OU", 2, 100, 1), ("BM", 1, 150, 1)]:
    dat, par, info = synth.make_problem(model, T, m, missing_frac=0.1, n_dim=nd)
    eng = Engine.from_data(dat)
    v, g = eng.eval(par, 1)
    ref = O.nllk(dat, par)
    print(model, v, ref, abs(v - ref) / abs(ref), "ms", eng.last_eval_ms, "launches", eng.last_eval_launches)
    eng.close()
