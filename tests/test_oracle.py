"""CPU tests that pin the oracle (numpy restatement + C restatement) against independent known
answers and against the committed golden fixtures.  No GPU needed."""
import warnings

import numpy as np
import pytest

import golden_util as G
import harness_util as H
from oracle import known_answers as KA
from oracle import oracle_c
from oracle import oracle_np as O
from smoothsde_b200 import synth


def grad_err(g, g_ref):
    scale = np.maximum(np.abs(g_ref), 1e-3 * np.max(np.abs(g_ref)))
    return np.max(np.abs(g - g_ref) / scale)


@pytest.fixture(scope="module")
def harness():
    return H.build_harness()


# ---------------------------------------------------------------------------------------------
# golden fixtures
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", G.names())
def test_numpy_oracle_reproduces_golden(name):
    dat, gold = G.load(name)
    v = O.nllk(dat, gold["par"])
    assert abs(v - gold["nllk"]) <= 1e-13 * abs(gold["nllk"])
    # fixtures carry their independent cross-checks
    assert abs(gold["known_answer"] - gold["nllk"]) <= 1e-12 * abs(gold["nllk"])
    if "nllk_mpmath" in gold:
        assert abs(gold["nllk_mpmath"] - gold["nllk"]) <= 1e-12 * abs(gold["nllk"])


@pytest.mark.parametrize("name", [n for n in G.names() if "_ssm_" not in n and "_userH_" not in n and "_decay_" not in n])     # the C oracle covers default-shaped BM, OU, CTCRW
def test_c_oracle_reproduces_golden(name):
    dat, gold = G.load(name)
    for threads in (1, 2):
        v, g = oracle_c.COracle(dat, nthreads=threads).eval(gold["par"], True)
        assert abs(v - gold["nllk"]) <= 1e-12 * abs(gold["nllk"])
        assert grad_err(g, gold["grad"]) <= 1e-9
    if dat["type"] == "CTCRW":
        aest = oracle_c.COracle(dat).aest(gold["par"])
        ID = dat["ID"]
        last = np.r_[ID[1:] != ID[:-1], True]
        assert np.max(np.abs(aest[~last] - gold["aest_all"][~last])) < 1e-10


# ---------------------------------------------------------------------------------------------
# known answers on fresh problems
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("model,T,m,miss,nd", [
    ("BM", 2, 150, 0.1, 1), ("BM", 1, 100, 0.0, 3), ("OU", 3, 100, 0.2, 1), ("OU", 1, 200, 0.0, 2),
    ("CTCRW", 2, 50, 0.2, 2), ("CTCRW", 4, 25, 0.0, 1),
])
def test_oracle_matches_known_answer(model, T, m, miss, nd):
    dat, par, _ = synth.make_problem(model, T, m, missing_frac=miss, n_dim=nd, seed=300 + m)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        v = O.nllk(dat, par)
    ka = KA.known_answer(dat, par)
    assert abs(v - ka) <= 1e-11 * abs(ka), (v, ka)


def test_oracle_matches_mpmath_ctcrw():
    dat, par, _ = synth.make_problem("CTCRW", 2, 30, missing_frac=0.1, n_dim=2, seed=77)
    par = par.copy()
    par[1:3] = [0.4, -0.1]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        v = O.nllk(dat, par)
    mp = KA.nllk_mpmath(dat, par)
    assert abs(v - mp) <= 1e-12 * abs(mp)


def test_ctcrw_cov_is_makeQ_with_swapped_order():
    """R/utility.R:188-196 (CTCRW_cov, (v, z) order) == makeQ_ctcrw (z, v), nllk_ctcrw.hpp:63-75."""
    from smoothsde_b200.simulate import ctcrw_cov
    beta, sigma, dt = 0.7, 1.3, 0.9
    qvv, qzz, qvz = ctcrw_cov(beta, sigma, dt)
    Q = O.makeQ_ctcrw(beta, sigma, dt, 1, float)
    assert np.allclose([Q[1, 1], Q[0, 0], Q[0, 1]], [qvv, qzz, qvz], rtol=1e-14)


def test_kalman_penalty_ignores_include_penalty_but_sde_honours_it():
    """include_penalty is only read by nllk_sde (nllk_sde.hpp:28,91)."""
    dat, par, _ = synth.make_problem("CTCRW", 1, 30, n_dim=1, seed=3)
    d0 = dict(dat, include_penalty=0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        assert O.nllk(dat, par) == O.nllk(d0, par)
    dat, par, _ = synth.make_problem("BM", 1, 30, n_dim=1, seed=3)
    d0 = dict(dat, include_penalty=0)
    p = O.split_par(dat, par)
    assert abs((O.nllk(dat, par) - O.nllk(d0, par)) - O.penalty_sde(dat, p["log_lambda"], p["coeff_re"])) < 1e-10


def test_first_and_last_row_of_each_track_have_zero_gradient():
    """parameters of a track's first and last row never reach the CTCRW nllk."""
    dat, par, _ = synth.make_problem("CTCRW", 2, 20, n_dim=2, seed=9)
    co = oracle_c.COracle(dat)
    co.eval(par, True)
    n = dat["obs"].shape[0]
    pb = co.par_bar.reshape(4, n)
    for r in (0, 19, 20, 39):
        assert np.all(pb[:, r] == 0.0)
    assert np.any(pb[:, 1] != 0.0)


# ---------------------------------------------------------------------------------------------
# C oracle vs numpy oracle + complex step on fresh problems, OpenMP on
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("model,T,m,miss,nd", [
    ("CTCRW", 3, 70, 0.15, 2), ("CTCRW", 5, 41, 0.3, 1), ("BM", 3, 80, 0.1, 2), ("OU", 4, 90, 0.1, 1),
])
def test_c_oracle_matches_numpy_and_complex_step(model, T, m, miss, nd):
    dat, par, _ = synth.make_problem(model, T, m, missing_frac=miss, n_dim=nd, seed=5)
    if model == "CTCRW":
        par = par.copy()
        par[1:1 + nd] = [0.3, -0.2][:nd]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        v = O.nllk(dat, par)
        g = O.grad_complex_step(dat, par)
    for threads in (1, 4):
        vc, gc = oracle_c.COracle(dat, nthreads=threads).eval(par, True)
        assert abs(v - vc) <= 1e-12 * abs(v)
        assert grad_err(gc, g) <= 1e-9


# ---------------------------------------------------------------------------------------------
# the engine's scan algebra (ctcrw_math.cuh compiled for the host) vs the oracle
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("T,m,miss,nd,lc,nt", [
    (3, 70, 0.15, 2, 8, 32), (1, 400, 0.05, 2, 4, 8), (6, 33, 0.3, 1, 8, 4), (2, 257, 0.0, 2, 16, 2),
    (40, 3, 0.0, 2, 8, 32),
])
def test_scan_algebra_matches_oracle(harness, T, m, miss, nd, lc, nt):
    dat, par, _ = synth.make_problem("CTCRW", T, m, missing_frac=miss, n_dim=nd, seed=21 + m)
    par = par.copy()
    par[1:1 + nd] = [0.3, -0.2][:nd]
    p = O.split_par(dat, par)
    co = oracle_c.COracle(dat)
    v, g = co.eval(par, True)
    pen = O.penalty_kalman(dat, p["log_lambda"], p["coeff_re"])
    eta = O.linear_predictor(dat, p["coeff_fe"], p["coeff_re"])
    n = dat["obs"].shape[0]
    pb_ref = co.par_bar.reshape(nd + 2, n).T
    for mode in (0, 1):
        llk, eb, gsig, _ = H.harness_ctcrw(harness, dat, eta, p["log_sigma_obs"], mode, lc=lc, nt=nt)
        assert abs((-llk + pen) - v) <= 1e-12 * abs(v)
        assert abs(gsig - g[0]) <= 1e-9 * max(abs(g[0]), 1e-3)
        assert np.max(np.abs(eb - pb_ref)) <= 1e-9 * max(np.max(np.abs(pb_ref)), 1.0)



# ---------------------------------------------------------------------------------------------
# tangent (Dual) instantiation of the same algebra = second-order adjoint: its directional
# derivative of the adjoint must match a Richardson-extrapolated central difference of the
# plain-double adjoint, sequentially (mode 0) and through the chunked scan (mode 1)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("T,m,miss,nd,lc,nt", [(3, 70, 0.15, 2, 8, 32), (2, 120, 0.1, 1, 4, 8), (5, 9, 0.0, 2, 8, 32)])
def test_tangent_algebra_matches_finite_differences(harness, T, m, miss, nd, lc, nt):
    dat, par, _ = synth.make_problem("CTCRW", T, m, missing_frac=miss, n_dim=nd, seed=5 + m)
    par = par.copy()
    par[1:1 + nd] = [0.3, -0.2][:nd]
    p = O.split_par(dat, par)
    eta = O.linear_predictor(dat, p["coeff_fe"], p["coeff_re"])
    rng = np.random.default_rng(3)
    eta_dot = rng.normal(size=eta.shape)
    lso, lso_dot = float(p["log_sigma_obs"]), 0.7

    def adjoint(t):
        _, eb, gsig, _ = H.harness_ctcrw(harness, dat, eta + t * eta_dot, lso + t * lso_dot, 0, lc=lc, nt=nt)
        return eb, gsig

    def richardson(f, k):
        d1 = (f(k) - f(-k)) / (2 * k)
        d2 = (f(k / 2) - f(-k / 2)) / k
        return (4 * d2 - d1) / 3

    k = 2e-3
    ebd_fd = richardson(lambda t: adjoint(t)[0], k)
    gsd_fd = richardson(lambda t: adjoint(t)[1], k)
    eb0, gs0 = adjoint(0.0)
    for mode in (0, 1):
        llk2, (eb, ebd), (g_lso, g_lso_dot) = H.harness_ctcrw_tangent(harness, dat, eta, eta_dot, lso, lso_dot, mode, lc=lc, nt=nt)
        assert np.max(np.abs(eb - eb0)) <= 1e-11 * max(np.max(np.abs(eb0)), 1.0)
        # d llk = -(grad . direction)
        dd = -(np.sum(eb0 * eta_dot) + gs0 * lso_dot)
        assert abs(llk2[1] - dd) <= 1e-9 * max(abs(dd), 1.0)
        scale = max(np.max(np.abs(ebd_fd)), 1.0)
        assert np.max(np.abs(ebd - ebd_fd)) <= 2e-7 * scale
        assert abs(g_lso_dot - gsd_fd) <= 2e-7 * max(abs(gsd_fd), 1.0)


# ---------------------------------------------------------------------------------------------
# one-state Kalman models (OU_SSM, BM_SSM): oracle vs dense MVN, engine algebra vs oracle
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("model,T,m,miss,nd", [("OU_SSM", 3, 40, 0.1, 2), ("OU_SSM", 2, 60, 0.0, 1), ("BM_SSM", 3, 40, 0.2, 2), ("BM_SSM", 1, 70, 0.1, 3)])
def test_ssm_oracle_matches_dense_mvn_and_engine_algebra(harness, model, T, m, miss, nd):
    from oracle import known_answers as KA
    dat, par, _ = synth.make_problem(model, T, m, missing_frac=miss, n_dim=nd, seed=31 + m, k=5)
    par = par.copy()
    par[1:1 + nd] = [0.3, -0.2, 0.1][:nd]
    v = O.nllk(dat, par)
    assert abs(KA.known_answer(dat, par) - v) <= 1e-11 * abs(v)
    g = O.grad_complex_step(dat, par)
    p = O.split_par(dat, par)
    pen = O.penalty_kalman(dat, p["log_lambda"], p["coeff_re"])
    eta = O.linear_predictor(dat, p["coeff_fe"], p["coeff_re"])
    import scipy.sparse as sp
    X = sp.hstack([sp.csr_matrix(dat["X_fe"]), sp.csr_matrix(dat["X_re"])], format="csr")
    for mode, lc, nt in ((0, 8, 32), (1, 4, 8), (1, 8, 2)):
        llk, eb, gsig = H.harness_kalman(harness, dat, eta, p["log_sigma_obs"], mode, lc=lc, nt=nt)
        assert abs((-llk + pen) - v) <= 1e-12 * abs(v)
        assert abs(gsig - g[0]) <= 1e-9 * max(abs(g[0]), 1e-3)
        gth = X.T @ eb.T.ravel()                      # d nllk / d [coeff_fe | coeff_re] without the penalty part
        p_fe = dat["X_fe"].shape[1]
        assert np.max(np.abs(gth[:p_fe] - g[1:1 + p_fe])) <= 1e-9 * max(1.0, np.max(np.abs(g)))
