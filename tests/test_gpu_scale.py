"""Direct CUDA-vs-oracle parity at 1e7+ rows (the 1e8-row tests compare the
engine with itself).  The device-built benchmark shapes are copied to the host and evaluated by the
C oracle (oracle/oracle_c.c -- pinned to the reference's own code by tests/test_ref.py) on all host
cores; tolerances are the north star's 1e-10 / 1e-7.  12 500 scan tiles per kernel, nearly all of
which take the constant-map look-back path; the same evaluation with that shortcut switched off
(ssde_debug_const_map_tol(-1)) must agree to rounding."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import oracle_c

pytestmark = pytest.mark.gpu


def grad_err(g, g_ref):
    scale = np.maximum(np.abs(g_ref), 1e-3 * np.max(np.abs(g_ref)))
    return np.max(np.abs(g - g_ref) / scale)


def host_oracle(info, T, m, k=10):
    """The device-built CTCRW problem as CSR arrays on the host (un-permuting the packed values)."""
    t = info["packed"]
    n, n_pad, lc, nnz, nv = info["n"], info["n_pad"], t["lc"], t["nnz_row"], t["nval_row"]
    val = t["val"].reshape(n_pad // (32 * lc), lc, nv, 32).permute(0, 3, 1, 2).reshape(n_pad, nv)[:n].cpu().numpy()
    if nv != nnz:        # aliased layout (design.cuh): a row stores [mu intercept | tau block]; mu2 = mu1's slot, nu = tau's
        assert nv == 1 + k and nnz == 2 + 2 * k
        val = np.concatenate([val[:, :1], val[:, :1], val[:, 1:], val[:, 1:]], axis=1)
    p_fe, p_re = info["p_fe"], info["p_re"]
    # rows j*n + i of [X_fe | X_re]: mu1 (1 nonzero), mu2 (1), tau (k), nu (k)
    cnt = np.concatenate([np.full(n, 1), np.full(n, 1), np.full(n, k), np.full(n, k)]).astype(np.int64)
    rowptr = np.concatenate([[0], np.cumsum(cnt)])
    tau_cols = np.array([2] + [p_fe + j for j in range(k - 1)], dtype=np.int32)
    nu_cols = np.array([3] + [p_fe + (k - 1) + j for j in range(k - 1)], dtype=np.int32)
    col = np.concatenate([np.zeros(n, np.int32), np.ones(n, np.int32), np.tile(tau_cols, n), np.tile(nu_cols, n)])
    data = np.concatenate([val[:, 0], val[:, 1], val[:, 2:2 + k].ravel(), val[:, 2 + k:].ravel()])
    del val
    ten = info["tensors"]
    cores = len(os.sched_getaffinity(0))
    return oracle_c.COracle.from_csr(
        "CTCRW", np.repeat(np.arange(1, T + 1), m).astype(float), ten["times"].cpu().numpy(), ten["obs"].cpu().numpy(),
        rowptr, col, data, p_fe, p_re, sp.csr_matrix(info["S"]), np.array([k - 1, k - 1]), a0=info["a0"],
        P0=np.diag([1.0, 10.0, 1.0, 10.0]), nthreads=cores)


@pytest.mark.parametrize("T,m,seed", [(128, 100000, 20260103),      # config[2] shape, 1.28e7 rows
                                      (512, 25000, 20260105)])      # config[4] shape, 1.28e7 rows
def test_device_built_shapes_match_the_c_oracle_at_1e7_rows(T, m, seed):
    from smoothsde_b200 import _lib, devgen
    eng, par, info = devgen.make_ctcrw_device(T, m, seed=seed, device=0)
    rng = np.random.default_rng(1)
    par = par.copy()
    par[1:3] = [0.3, -0.2]                      # B mu
    par[3:5] = [0.2, -0.1]
    par[5:7] = [0.5, -0.5]                      # log lambda
    par[7:] = 0.2 * rng.standard_normal(par.size - 7)
    v, g = eng.eval(par, order=1)
    co = host_oracle(info, T, m)
    ref_v, ref_g = co.eval(par, True)
    assert abs(v - ref_v) <= 1e-10 * abs(ref_v), (v, ref_v)
    assert grad_err(g, ref_g) <= 1e-7, (g, ref_g)
    # A/B: the look-back without the constant-map shortcut
    lib = _lib.load()
    assert lib.ssde_debug_const_map_tol(0, -1.0) == 0
    try:
        v2, g2 = eng.eval(par, order=1)
    finally:
        assert lib.ssde_debug_const_map_tol(0, 1e-60) == 0
    assert abs(v2 - v) <= 1e-13 * abs(v), (v, v2)
    assert grad_err(g2, g) <= 1e-11
    v3, g3 = eng.eval(par, order=1)             # shortcut back on: same result as before
    assert v3 == v
    eng.close()


def test_aliased_and_plain_design_layouts_agree_at_1e7_rows():
    """tau and nu share one smooth of time: the default layout stores their values once per row (alias
    flags, 11 of 22 doubles).  Same data, both layouts, through both kernels: the linear predictors add
    the same products in the same order, so nllk and gradient agree to the last bits."""
    from smoothsde_b200 import devgen
    T, m = 128, 100000
    res = []
    for alias in (True, False):
        eng, par, info = devgen.make_ctcrw_device(T, m, seed=20260107, device=0, alias=alias)
        assert info["packed"]["nval_row"] == (11 if alias else 22)
        par = par.copy()
        par[1:5] = [0.3, -0.2, 0.2, -0.1]
        res.append(eng.eval(par, order=1))
        eng.close()
        del eng, info
    (v1, g1), (v0, g0) = res
    assert abs(v1 - v0) <= 1e-13 * abs(v0), (v1, v0)
    assert grad_err(g1, g0) <= 1e-11, (g1, g0)


def test_device_built_ou_with_random_intercepts_matches_the_c_oracle_at_1e7_rows():
    """The OU half of configs[4] / configs[1]: mu, tau ~ s(time, k = 10) + s(ID, bs = "re"), kappa ~ 1,
    512 tracks x 25 000 steps (1.28e7 rows, p_re = 1042; warp-tiles that straddle two tracks carry both
    tracks' random-intercept columns).  The oracle sees the natural (row, column, value) design."""
    from smoothsde_b200 import devgen
    T, m, k = 512, 25000, 10
    eng, par, info = devgen.make_ou_device(T, m, seed=20260102, device=0)
    n, p_fe, p_re = info["n"], info["p_fe"], info["p_re"]
    rng = np.random.default_rng(2)
    par = par.copy()
    par[:3] = [0.2, 0.1, np.log(1.3)]
    par[3:7] = [0.3, -0.4, 0.2, 0.5]            # log lambda
    par[7:] = 0.2 * rng.standard_normal(p_re)
    v, g = eng.eval(par, order=1)
    # natural design: rows j*n + i; mu: [intercept, 9 spline, own random intercept], tau alike, kappa: [intercept]
    Bz = info["Bz1"].cpu().numpy()                                  # [m, k-1], the same for every track
    km1 = k - 1
    trk = np.repeat(np.arange(T), m)
    cnt = np.concatenate([np.full(n, 2 + km1), np.full(n, 2 + km1), np.full(n, 1)]).astype(np.int64)
    rowptr = np.concatenate([[0], np.cumsum(cnt)])
    spl = np.arange(km1)
    col_mu = np.empty((n, 2 + km1), np.int32)
    col_mu[:, 0] = 0
    col_mu[:, 1:1 + km1] = p_fe + spl
    col_mu[:, 1 + km1] = p_fe + km1 + trk
    col_tau = np.empty((n, 2 + km1), np.int32)
    col_tau[:, 0] = 1
    col_tau[:, 1:1 + km1] = p_fe + km1 + T + spl
    col_tau[:, 1 + km1] = p_fe + 2 * km1 + T + trk
    vals = np.empty((n, 2 + km1))
    vals[:, 0] = 1.0
    vals[:, 1:1 + km1] = np.tile(Bz, (T, 1))
    vals[:, 1 + km1] = 1.0
    col = np.concatenate([col_mu.ravel(), col_tau.ravel(), np.full(n, 2, np.int32)])
    data = np.concatenate([vals.ravel(), vals.ravel(), np.ones(n)])
    ten = info["tensors"]
    co = oracle_c.COracle.from_csr("OU", trk.astype(float) + 1, ten["times"].cpu().numpy(), ten["obs"].cpu().numpy().reshape(n, 1),
                                   rowptr, col, data, p_fe, p_re, sp.csr_matrix(info["S"]), np.array([km1, T, km1, T]),
                                   nthreads=len(os.sched_getaffinity(0)))
    ref_v, ref_g = co.eval(par, True)
    assert abs(v - ref_v) <= 1e-10 * abs(ref_v), (v, ref_v)
    assert grad_err(g, ref_g) <= 1e-7
    eng.close()
