"""Decay terms of nllk_sde (BM / OU): nllk_sde.hpp:31-33,47-59, R/sde.R:162-180,303-326,635-649.
Columns col_decay of X_re are multiplied row by row by exp(-exp(log_decay[ind_decay]) * t_decay);
log_decay sits between log_lambda and coeff_re in the parameter vector (nllk_sde.hpp:42-45)."""
import warnings

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import oracle_np as O
from smoothsde_b200 import synth


def decay_problem(model, T, m, nd, seed, n_dec=2):
    dat, par, info = synth.make_problem(model, T, m, missing_frac=0.1, n_dim=nd, seed=seed, k=5)
    n = dat["obs"].shape[0]
    n_par = dat["X_fe"].shape[0] // n
    t = np.asarray(dat["times"], float)
    ID = np.asarray(dat["ID"])
    first = np.r_[True, ID[1:] != ID[:-1]]
    t0 = t[np.maximum.accumulate(np.where(first, np.arange(n), 0))]     # time of the track's first row
    dat["t_decay"] = np.tile((t - t0) * 0.02, n_par) * np.repeat(1.0 + 0.1 * np.arange(n_par), n)
    p_re = dat["X_re"].shape[1]
    cols = np.arange(1, min(p_re, 6) + 1)
    dat["col_decay"] = cols
    dat["ind_decay"] = 1 + (np.arange(cols.size) % n_dec)
    n_s = np.atleast_1d(dat["ncol_re"]).size
    o = info["p_fe"] + n_s
    rng = np.random.default_rng(seed)
    full = np.r_[par[:o], rng.normal(0.0, 0.4, n_dec), par[o:]]
    return dat, full, info


@pytest.mark.parametrize("model,nd", [("BM", 2), ("OU", 1)])
def test_oracle_decay_equals_prescaled_design(model, nd):
    dat, full, info = decay_problem(model, 3, 50, nd, 3)
    p = O.split_par(dat, full)
    assert p["log_decay"].size == 2 and p["coeff_re"].size == dat["X_re"].shape[1]
    v = O.nllk(dat, full)
    X = sp.csr_matrix(dat["X_re"]).toarray()
    for c, k in zip(dat["col_decay"], dat["ind_decay"]):
        X[:, c - 1] *= np.exp(-np.exp(p["log_decay"][k - 1]) * dat["t_decay"])
    plain = {k_: v_ for k_, v_ in dat.items() if k_ not in ("t_decay", "col_decay", "ind_decay")}
    plain["X_re"] = sp.csr_matrix(X)
    v2 = O.nllk(plain, np.r_[p["coeff_fe"], p["log_lambda"], p["coeff_re"]])
    assert abs(v - v2) <= 1e-13 * abs(v)
    # the gradient w.r.t. log_decay is not zero (the term matters)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        g = O.grad_complex_step(dat, full)
    o = info["p_fe"] + np.atleast_1d(dat["ncol_re"]).size
    assert np.all(np.abs(g[o:o + 2]) > 1e-8)


def test_sde_host_layer_builds_the_decay_lists():
    from smoothsde_b200.sde import SDE
    rng = np.random.default_rng(1)
    n = 120
    t = np.arange(n, dtype=float)
    data = {"ID": np.ones(n, int), "time": t, "z": np.cumsum(rng.normal(size=n))}
    forms = {"mu": "~s(time, k=5, bs='cs')", "sigma": "~1"}
    with pytest.raises(ValueError, match="number of parameters"):
        SDE(formulas=forms, data=data, type="BM", response="z", other_data={"t_decay": t, "decay_term": "mu.s(time)", "ind_decay": [1] * 4})
    sde = SDE(formulas=forms, data=data, type="BM", response="z",
              other_data={"t_decay": np.tile(t * 0.05, 2), "decay_term": "mu.s(time)", "ind_decay": [1, 1, 1, 1]})
    assert list(sde.other_data()["col_decay"]) == [1, 2, 3, 4] and sde.rho().tolist() == [1.0]
    tmb_dat, tmb_par, map_, random = sde.tmb_lists()
    assert list(tmb_par.keys()) == ["coeff_fe", "log_lambda", "log_decay", "coeff_re"]      # nllk_sde.hpp:42-45
    assert "log_decay" not in map_ and tmb_dat["t_decay"].size == 2 * n
    Xd = sde.X_re_decay().toarray()
    X0 = sp.csr_matrix(sde.mats().X_re).toarray()
    assert np.allclose(Xd[:, :4], X0[:, :4] * np.exp(-1.0 * np.tile(t * 0.05, 2))[:, None], rtol=1e-14, atol=0)
    # without decay terms the entry exists but is mapped off (R/sde.R:645-648)
    sde0 = SDE(formulas=forms, data=data, type="BM", response="z")
    _, tmb_par0, map0, _ = sde0.tmb_lists()
    assert "log_decay" in tmb_par0 and map0["log_decay"] == [None]
    with pytest.raises(ValueError, match="no decaying terms"):
        sde0.X_re_decay()


# ---------------------------------------------------------------------------------------------
# GPU parity
# ---------------------------------------------------------------------------------------------
def grad_err(g, g_ref):
    scale = np.maximum(np.abs(g_ref), 1e-3 * np.max(np.abs(g_ref)))
    return np.max(np.abs(g - g_ref) / scale)


@pytest.mark.gpu
@pytest.mark.parametrize("model,T,m,nd,n_dec", [("BM", 3, 70, 2, 2), ("OU", 4, 90, 1, 2), ("OU", 2, 700, 2, 3), ("BM", 1, 1300, 1, 1)])
def test_decay_engine_matches_oracle(model, T, m, nd, n_dec):
    from smoothsde_b200.engine import Engine
    dat, full, info = decay_problem(model, T, m, nd, 5 + m, n_dec)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref, g_ref = O.nllk(dat, full), O.grad_complex_step(dat, full)
    eng = Engine.from_data(dat)
    assert eng.layout["log_decay"] == (info["p_fe"] + np.atleast_1d(dat["ncol_re"]).size, n_dec)
    assert eng.n_par == full.size
    v0, _ = eng.eval(full, order=0)
    v, g = eng.eval(full, order=1)
    assert abs(v0 - v) <= 1e-13 * abs(v)
    assert abs(v - ref) <= 1e-10 * abs(ref), (v, ref)
    assert grad_err(g, g_ref) <= 1e-7, (g, g_ref)
    # Hessian-vector products (tangent pass through the decay factors)
    rng = np.random.default_rng(2)
    d = rng.normal(size=full.size)
    _, _, hv = eng.hvp(full, d)
    k = 1e-3
    gr = lambda p: O.grad_complex_step(dat, p)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        d1 = (gr(full + k * d) - gr(full - k * d)) / (2 * k)
        d2 = (gr(full + 0.5 * k * d) - gr(full - 0.5 * k * d)) / k
    h_ref = (4 * d2 - d1) / 3
    scale = np.maximum(np.abs(h_ref), 1e-3 * np.max(np.abs(h_ref)))
    assert np.max(np.abs(hv.ravel() - h_ref) / scale) <= 1e-6
    eng.close()


@pytest.mark.gpu
def test_decay_fit_updates_rho():
    """SDE$fit() of a decay model estimates log_decay with the other outer parameters and stores
    exp(log_decay) as rho (R/sde.R:715-719)."""
    from smoothsde_b200.sde import SDE
    rng = np.random.default_rng(7)
    n = 400
    t = np.arange(n, dtype=float)
    mu_t = 0.8 * np.exp(-0.03 * t) * np.sin(t / 15.0)
    z = np.cumsum(mu_t + 0.3 * rng.normal(size=n))
    data = {"ID": np.ones(n, int), "time": t, "z": z}
    sde = SDE(formulas={"mu": "~s(time, k=6, bs='cs')", "sigma": "~1"}, data=data, type="BM", response="z",
              other_data={"t_decay": np.tile(t / n, 2), "decay_term": "mu.s(time)", "ind_decay": [1] * 5})
    res = sde.fit(maxiter=60)
    assert np.isfinite(res.fun)
    assert "log_decay" in sde.tmb_obj().engine.layout
    off, size = sde.tmb_obj().engine.layout["log_decay"]
    assert abs(sde.rho()[0] - np.exp(sde._par_all[off])) < 1e-12
    # the marginal gradient at the optimum is small
    assert np.max(np.abs(sde.tmb_obj().gr(res.x))) < 1e-2
