"""GPU parity against the REFERENCE ITSELF: the CUDA engine (through the C ABI) vs

  * the committed outputs of the reference's own objective (`ref_*` in tests/golden/*.npz,
    tests/golden/make_ref_golden.py), and
  * the prebuilt oracle/_ref/libsmoothsde_ref.so (the reference's unmodified templates compiled
    against oracle/tmb_shim/TMB.hpp) on freshly seeded problems of a few thousand rows.

Tolerances are the north star's: nllk 1e-10, gradient 1e-7 relative; Hessian 1e-6."""
import numpy as np
import pytest

import golden_util as G
from oracle import oracle_ref
from smoothsde_b200 import synth
from smoothsde_b200.engine import Engine

pytestmark = pytest.mark.gpu


def grad_err(g, g_ref):
    scale = np.maximum(np.abs(g_ref), 1e-3 * np.max(np.abs(g_ref)))
    return np.max(np.abs(g - g_ref) / scale)


@pytest.mark.parametrize("name", G.names())
def test_engine_reproduces_reference_vectors(name):
    dat, gold = G.load(name)
    eng = Engine.from_data(dat)
    v, g, H = eng.hessian(gold["par"])
    assert abs(v - gold["ref_nllk"]) <= 1e-10 * abs(gold["ref_nllk"]), (v, gold["ref_nllk"])
    assert grad_err(g, gold["ref_grad"]) <= 1e-7
    assert grad_err(H, gold["ref_hess"]) <= 1e-6
    if "ref_aest_all" in gold:
        eng.eval(gold["par"], order=0)
        n, ns = gold["ref_aest_all"].shape
        aest = eng.report(n, dat["obs"].shape[1], ns)
        ID = dat["ID"]
        last = np.r_[ID[1:] != ID[:-1], True]     # REPORT rows that end a track hold a cross-track prediction
        assert np.max(np.abs(aest[~last] - gold["ref_aest_all"][~last])) <= 1e-9 * max(1.0, np.abs(gold["ref_aest_all"]).max())
    eng.close()


LIVE = [
    ("CTCRW", 3, 2500, 0.05, 2),      # several scan tiles of both kernels
    ("CTCRW", 40, 30, 0.2, 1),
    ("OU", 5, 800, 0.1, 1),
    ("BM", 2, 1500, 0.05, 2),
    ("OU_SSM", 2, 1500, 0.05, 2),
    ("BM_SSM", 3, 700, 0.1, 3),
]


@pytest.mark.skipif(not oracle_ref.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("model,T,m,miss,nd", LIVE)
def test_engine_matches_live_reference(model, T, m, miss, nd):
    dat, par, info = synth.make_problem(model, T, m, missing_frac=miss, n_dim=nd, seed=500 + T + m)
    if model in ("CTCRW", "OU_SSM", "BM_SSM"):
        par = par.copy()
        par[1:1 + nd] = [0.3, -0.2, 0.1][:nd]
    R = oracle_ref.RefOracle(dat)
    rv, rg = R.eval(par)
    eng = Engine.from_data(dat)
    v, g = eng.eval(par, order=1)
    assert abs(v - rv) <= 1e-10 * abs(rv), (v, rv)
    assert grad_err(g, rg) <= 1e-7
    rng = np.random.default_rng(3)
    d = rng.standard_normal(par.size)
    _, _, hv = eng.hvp(par, d[:, None])
    _, _, rhv = R.hvp(par, d)
    assert grad_err(hv[:, 0], rhv) <= 1e-6
    eng.close()
