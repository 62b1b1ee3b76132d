"""GPU parity tests: CUDA engine (through the C ABI) vs the CPU oracle on identical inputs.

Tolerances are the north star's: nllk 1e-10 relative, gradient components 1e-7 relative
(relative to max(|g|, 1e-3 * max|g|) so that exact zeros do not blow the ratio up)."""
import numpy as np
import pytest

from oracle import oracle_np as O
from smoothsde_b200 import synth
from smoothsde_b200.engine import Engine

pytestmark = pytest.mark.gpu

NLLK_RTOL = 1e-10
GRAD_RTOL = 1e-7


def grad_err(g, g_ref):
    scale = np.maximum(np.abs(g_ref), 1e-3 * np.max(np.abs(g_ref)))
    return np.max(np.abs(g - g_ref) / scale)


CASES = [
    ("CTCRW", 3, 70, 0.15, 2),
    ("CTCRW", 1, 300, 0.0, 2),
    ("CTCRW", 5, 41, 0.3, 1),
    ("CTCRW", 2, 1500, 0.05, 2),     # spans several scan tiles
    ("BM", 1, 200, 0.1, 1),
    ("BM", 3, 80, 0.1, 2),
    ("OU", 4, 90, 0.1, 1),
    ("OU", 2, 120, 0.0, 2),
]


@pytest.mark.parametrize("model,T,m,miss,nd", CASES)
def test_nllk_and_gradient_match_oracle(model, T, m, miss, nd):
    dat, par, info = synth.make_problem(model, T, m, missing_frac=miss, n_dim=nd, seed=7 + T + m)
    if model == "CTCRW":
        par = par.copy()
        par[1:1 + nd] = [0.3, -0.2][:nd]          # exercise B*mu
    ref = O.nllk(dat, par)
    eng = Engine.from_data(dat)
    v0, _ = eng.eval(par, order=0)
    v, g = eng.eval(par, order=1)
    assert v0 == v
    assert abs(v - ref) <= NLLK_RTOL * abs(ref), (v, ref)
    g_ref = O.grad_complex_step(dat, par)
    assert grad_err(g, g_ref) <= GRAD_RTOL, (g, g_ref)
    # evaluating twice gives the same value (epoch / ticket reset)
    v2, g2 = eng.eval(par, order=1)
    assert abs(v2 - v) <= 1e-13 * abs(v)
    assert grad_err(g2, g) <= 1e-12
    eng.close()


def test_report_aest_matches_oracle():
    dat, par, info = synth.make_problem("CTCRW", 3, 60, missing_frac=0.1, n_dim=2)
    p = O.split_par(dat, par)
    _, aest_ref = O.nllk_ctcrw(dat, **p, return_aest=True)
    eng = Engine.from_data(dat)
    eng.eval(par, order=0)
    aest = eng.report(dat["obs"].shape[0], 2)
    ID = dat["ID"]
    last = np.r_[ID[1:] != ID[:-1], True]     # the reference predicts across tracks there (garbage)
    assert np.max(np.abs(aest[~last] - aest_ref[~last])) < 1e-9
    eng.close()
