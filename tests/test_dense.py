"""Coupled ("dense") Kalman filter: user-supplied H_array (nllk_ctcrw.hpp:203-205, R/sde.R:593-598)
and P0 matrices that are not of the default shape.  CPU part: the host-compiled scan algebra of
DenseModel (models.cuh / dense_math.cuh) against the numpy oracle (value, complex-step gradient,
REPORT(aest_all)), sequentially and through the emulated chunked scan."""
import warnings

import numpy as np
import pytest
import scipy.sparse as sp

import harness_util as H
from oracle import oracle_np as O
from smoothsde_b200 import synth


@pytest.fixture(scope="module")
def harness():
    return H.build_harness()


def random_spd(rng, m, scale=1.0):
    A = rng.normal(size=(m, m))
    return scale * (A @ A.T / m + 0.5 * np.eye(m))


def dense_problem(model, T, m, nd, miss, seed, user_H, user_P0):
    dat, par, info = synth.make_problem(model, T, m, missing_frac=miss, n_dim=nd, seed=seed, k=5)
    par = par.copy()
    par[1:1 + nd] = [0.3, -0.2, 0.1][:nd]
    rng = np.random.default_rng(seed + 100)
    if user_P0:
        dat["P0"] = random_spd(rng, dat["P0"].shape[0], 3.0)
    if user_H:
        n = dat["obs"].shape[0]
        dat["H_array"] = np.stack([random_spd(rng, nd, 0.02) for _ in range(n)], axis=2)
    return dat, par


CASES = [
    # model, T, m, nd, miss, user_H, user_P0
    ("CTCRW", 3, 50, 2, 0.15, True, False),
    ("CTCRW", 2, 70, 2, 0.1, False, True),
    ("CTCRW", 2, 40, 2, 0.0, True, True),
    ("CTCRW", 2, 60, 1, 0.1, True, True),
    ("OU_SSM", 3, 40, 2, 0.1, True, True),
    ("BM_SSM", 2, 40, 3, 0.2, True, False),
    ("BM_SSM", 2, 50, 2, 0.1, False, True),
    ("CTCRW", 3, 50, 2, 0.15, False, False),       # default shapes through the coupled algebra
]


@pytest.mark.parametrize("model,T,m,nd,miss,user_H,user_P0", CASES)
def test_dense_algebra_matches_oracle(harness, model, T, m, nd, miss, user_H, user_P0):
    dat, par = dense_problem(model, T, m, nd, miss, 41 + m, user_H, user_P0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        v = O.nllk(dat, par)
        g = O.grad_complex_step(dat, par)
    p = O.split_par(dat, par)
    pen = O.penalty_kalman(dat, p["log_lambda"], p["coeff_re"])
    eta = O.linear_predictor(dat, p["coeff_fe"], p["coeff_re"])
    X = sp.hstack([sp.csr_matrix(dat["X_fe"]), sp.csr_matrix(dat["X_re"])], format="csr")
    p_fe = dat["X_fe"].shape[1]
    aest_ref = None
    for mode, lc, nt in ((0, 8, 64), (1, 8, 4), (1, 4, 2), (1, 16, 64)):
        llk, eb, gsig, aest = H.harness_dense(harness, dat, eta, p["log_sigma_obs"], mode, lc=lc, nt=nt, want_aest=True)
        assert abs((-llk + pen) - v) <= 1e-11 * abs(v), (mode, lc, nt)
        gth = X.T @ eb.T.ravel()
        assert np.max(np.abs(gth[:p_fe] - g[1:1 + p_fe])) <= 1e-8 * max(1.0, np.max(np.abs(g))), (mode, lc, nt)
        # with a user H the objective does not depend on log_sigma_obs (R/sde.R:593-595 maps it off)
        assert abs(gsig - g[0]) <= 1e-8 * max(abs(g[0]), 1e-3)
        if user_H:
            assert gsig == 0.0
        if aest_ref is None:
            aest_ref = aest
        assert np.max(np.abs(aest - aest_ref)) <= 1e-9 * max(1.0, np.max(np.abs(aest_ref)))
    fn = {"CTCRW": O.nllk_ctcrw, "OU_SSM": O.nllk_ou_ssm, "BM_SSM": O.nllk_bm_ssm}[model]
    _, aest_o = fn(dat, p["log_sigma_obs"], p["coeff_fe"], p["log_lambda"], p["coeff_re"], return_aest=True)
    # rows that end a track hold a prediction across the track boundary (discarded by both sides)
    ID = np.asarray(dat["ID"])
    last = np.r_[ID[1:] != ID[:-1], True]
    aest_o = np.asarray(aest_o, float)
    assert np.max(np.abs(aest_ref[~last] - aest_o[~last])) <= 1e-9 * max(1.0, np.max(np.abs(aest_o[~last])))


@pytest.mark.parametrize("model,T,m,nd,user_H", [("CTCRW", 2, 40, 2, True), ("CTCRW", 2, 40, 2, False), ("OU_SSM", 2, 30, 2, True)])
def test_dense_tangent_matches_finite_differences(harness, model, T, m, nd, user_H):
    dat, par = dense_problem(model, T, m, nd, 0.1, 7 + m, user_H, True)
    p = O.split_par(dat, par)
    eta = O.linear_predictor(dat, p["coeff_fe"], p["coeff_re"])
    rng = np.random.default_rng(3)
    eta_dot = rng.normal(size=eta.shape)
    lso, lso_dot = float(p["log_sigma_obs"]), 0.7

    def adjoint(t):
        _, eb, gsig, _ = H.harness_dense(harness, dat, eta + t * eta_dot, lso + t * lso_dot, 0)
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
        llk2, (eb, ebd), (g_lso, g_lso_dot) = H.harness_dense(harness, dat, eta, lso, mode, lc=4, nt=8, eta_dot=eta_dot, lso_dot=lso_dot)
        assert np.max(np.abs(eb - eb0)) <= 1e-10 * max(np.max(np.abs(eb0)), 1.0)
        dd = -(np.sum(eb0 * eta_dot) + gs0 * lso_dot)
        assert abs(llk2[1] - dd) <= 1e-8 * max(abs(dd), 1.0)
        assert np.max(np.abs(ebd - ebd_fd)) <= 5e-7 * max(np.max(np.abs(ebd_fd)), 1.0)
        assert abs(g_lso_dot - gsd_fd) <= 5e-7 * max(abs(gsd_fd), 1.0)


@pytest.mark.parametrize("nd,user_H,user_P0", [(2, True, True), (2, True, False), (1, True, True), (2, False, True)])
def test_oracle_with_user_H_and_P0_matches_dense_mvn(nd, user_H, user_P0):
    """Pins the oracle's H_array / general-P0 branches without any recursion: one multivariate
    normal density per track (oracle/known_answers.py)."""
    from oracle import known_answers as KA
    dat, par = dense_problem("CTCRW", 2, 25, nd, 0.15, 61, user_H, user_P0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        v = O.nllk(dat, par)
    assert abs(KA.known_ctcrw(dat, par) - v) <= 1e-10 * abs(v)


def test_pivoted_inverse_handles_vanishing_leading_minors(harness):
    """(I + C J) of the scan combine has positive eigenvalues but is not symmetric: with
    C = [[1, 3], [3, 10]], J = [[1, -3], [-3, 10]] its (1, 1) entry is 1 + 1 - 9 = -7, and leading
    minors can vanish altogether -- elimination without row exchanges would divide by zero."""
    import ctypes
    rng = np.random.default_rng(0)
    mats = [np.array([[0.0, 1.0], [1.0, 0.0]]), np.array([[0.0, 2.0, 0.0], [0.0, 0.0, 3.0], [4.0, 0.0, 0.0]]),
            np.eye(2) + np.array([[1.0, 3.0], [3.0, 10.0]]) @ np.array([[1.0, -3.0], [-3.0, 10.0]])]
    for n in (1, 2, 3, 4):
        for _ in range(20):
            A, B = rng.normal(size=(n, n)), rng.normal(size=(n, n))
            mats.append(np.eye(n) + (A @ A.T) @ (B @ B.T))
    P = ctypes.POINTER(ctypes.c_double)
    for X in mats:
        X = np.ascontiguousarray(X, float)
        Xi = np.zeros_like(X)
        assert harness.harness_inv_general(X.shape[0], X.ctypes.data_as(P), Xi.ctypes.data_as(P)) == 0
        ref = np.linalg.inv(X)
        assert np.max(np.abs(Xi - ref)) <= 1e-12 * np.linalg.cond(X) * np.max(np.abs(ref))
