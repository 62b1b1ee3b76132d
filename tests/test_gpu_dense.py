"""GPU parity of the coupled Kalman filter (user H_array, nllk_ctcrw.hpp:203-205 / R/sde.R:593-598,
and P0 matrices that are not of the default shape) against the numpy oracle: nllk 1e-10, gradient
1e-7, Hessian-vector products 1e-6, REPORT(aest_all), time shards, the SDE host layer."""
import warnings

import numpy as np
import pytest

from oracle import oracle_np as O
from smoothsde_b200 import _lib as L
from smoothsde_b200.engine import Engine
from test_dense import dense_problem

pytestmark = pytest.mark.gpu

NLLK_RTOL = 1e-10
GRAD_RTOL = 1e-7
HESS_RTOL = 1e-6


def grad_err(g, g_ref):
    scale = np.maximum(np.abs(g_ref), 1e-3 * np.max(np.abs(g_ref)))
    return np.max(np.abs(g - g_ref) / scale)


def oracle(dat, par, grad=True):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return O.nllk(dat, par), (O.grad_complex_step(dat, par) if grad else None)


CASES = [
    # model, T, m, nd, miss, user_H, user_P0
    ("CTCRW", 3, 70, 2, 0.15, True, False),
    ("CTCRW", 2, 1300, 2, 0.05, True, True),       # several scan tiles (512 rows each)
    ("CTCRW", 40, 3, 2, 0.0, True, False),         # many tiny tracks
    ("CTCRW", 2, 300, 2, 0.1, False, True),        # H = sigma_obs^2 I with a coupled P0: d/d log_sigma_obs
    ("CTCRW", 3, 200, 1, 0.1, True, True),
    ("OU_SSM", 3, 400, 2, 0.1, True, True),
    ("BM_SSM", 2, 300, 3, 0.2, True, False),
    ("BM_SSM", 2, 700, 2, 0.1, False, True),
    ("OU_SSM", 2, 300, 1, 0.1, True, False),       # n_dim = 1: a scalar h_i per row
    ("CTCRW", 1, 2, 2, 0.0, True, False),          # a single transition-free track (two rows)
    ("CTCRW", 7, 301, 2, 0.5, True, True),         # half of the rows missing
    ("CTCRW", 2, 512, 2, 0.02, True, False),       # track boundary exactly on a tile boundary (64 threads x 8 rows)
    ("BM_SSM", 30, 4, 3, 0.2, True, True),         # many tiny tracks, three coupled dimensions
]


@pytest.mark.parametrize("model,T,m,nd,miss,user_H,user_P0", CASES)
def test_coupled_filter_matches_oracle(model, T, m, nd, miss, user_H, user_P0):
    dat, par = dense_problem(model, T, m, nd, miss, 3 + T + m, user_H, user_P0)
    ref, g_ref = oracle(dat, par)
    eng = Engine.from_data(dat)
    v0, _ = eng.eval(par, order=0)
    v, g = eng.eval(par, order=1)
    assert abs(v0 - v) <= 1e-14 * abs(v), (v0, v)           # forward re-run (order 0) vs the adjoint kernel's recomputation
    assert abs(v - ref) <= NLLK_RTOL * abs(ref), (v, ref)
    assert grad_err(g, g_ref) <= GRAD_RTOL, (g, g_ref)
    if user_H:
        assert g[0] == 0.0                                  # log_sigma_obs does not enter (R/sde.R:593-595)
    v2, g2 = eng.eval(par, order=1)
    assert abs(v2 - v) <= 1e-13 * abs(v) and grad_err(g2, g) <= 1e-12
    eng.close()


def test_default_shapes_through_the_coupled_filter_equal_the_decoupled_engine():
    """A P0 that differs from the default shape by nothing the likelihood can see at 1e-10: both
    code paths must agree with each other far below the tolerance."""
    dat, par = dense_problem("CTCRW", 3, 900, 2, 0.1, 5, False, False)
    e1 = Engine.from_data(dat)
    v1, g1 = e1.eval(par, order=1)
    dat2 = dict(dat, P0=np.asarray(dat["P0"], float).copy())
    dat2["P0"][0, 2] = dat2["P0"][2, 0] = 1e-300            # not block-diagonal any more -> coupled filter
    e2 = Engine.from_data(dat2)
    v2, g2 = e2.eval(par, order=1)
    assert abs(v1 - v2) <= 1e-12 * abs(v1)
    assert grad_err(g2, g1) <= 1e-9
    e1.close(); e2.close()


def test_coupled_report_aest_matches_oracle():
    dat, par = dense_problem("CTCRW", 3, 700, 2, 0.1, 11, True, True)
    p = O.split_par(dat, par)
    _, aest_ref = O.nllk_ctcrw(dat, **p, return_aest=True)
    eng = Engine.from_data(dat)
    eng.eval(par, order=0)
    aest = eng.report(dat["obs"].shape[0], 2)
    ID = np.asarray(dat["ID"])
    last = np.r_[ID[1:] != ID[:-1], True]
    assert np.max(np.abs(aest[~last] - np.asarray(aest_ref, float)[~last])) < 1e-9
    eng.close()


@pytest.mark.parametrize("model,T,m,nd,user_H", [("CTCRW", 2, 600, 2, True), ("CTCRW", 2, 150, 2, False), ("OU_SSM", 2, 200, 2, True)])
def test_coupled_hessian_vector_products_match_oracle_differences(model, T, m, nd, user_H):
    dat, par = dense_problem(model, T, m, nd, 0.1, 17 + m, user_H, True)
    rng = np.random.default_rng(2)
    dirs = rng.normal(size=(par.size, 2))
    dirs[:, 1] = 0.0
    dirs[-1, 1] = 1.0
    eng = Engine.from_data(dat)
    v, g, hv = eng.hvp(par, dirs)
    ref_v, ref_g = oracle(dat, par)
    assert abs(v - ref_v) <= NLLK_RTOL * abs(ref_v)
    assert grad_err(g, ref_g) <= GRAD_RTOL
    k = 1e-3
    gr = lambda p: oracle(dat, p)[1]
    for c in range(dirs.shape[1]):
        d = dirs[:, c]
        d1 = (gr(par + k * d) - gr(par - k * d)) / (2 * k)
        d2 = (gr(par + 0.5 * k * d) - gr(par - 0.5 * k * d)) / k
        ref = (4 * d2 - d1) / 3
        scale = np.maximum(np.abs(ref), 1e-3 * np.max(np.abs(ref)))
        assert np.max(np.abs(hv[:, c] - ref) / scale) <= HESS_RTOL, (c, np.max(np.abs(hv[:, c] - ref)))
    eng.close()


def test_coupled_time_and_track_shards_match_the_single_handle():
    from smoothsde_b200 import sharded as S
    dat, par = dense_problem("CTCRW", 1, 5000, 2, 0.05, 23, True, True)
    e1 = Engine.from_data(dat)
    ref_v, ref_g = e1.eval(par, 1)
    ts = S.TimeShardedEngine(dat, devices=[0, 0, 0])
    v, g = ts.eval(par)
    assert abs(v - ref_v) <= 1e-11 * abs(ref_v), (v, ref_v)
    assert grad_err(g, ref_g) <= 1e-9
    ts.close(); e1.close()
    dat, par = dense_problem("CTCRW", 6, 400, 2, 0.05, 29, True, False)
    e1 = Engine.from_data(dat)
    ref_v, ref_g = e1.eval(par, 1)
    # two track shards (what two ranks hold); the second leaves the penalty to the first
    v, g = 0.0, 0.0
    for r, (lo, hi) in enumerate(S.split_tracks(dat["ID"], 2)):
        sub, cp, cn, _ = S.shard_rows(dat, lo, hi)
        assert not cp and not cn and sub["H_array"].shape[2] == hi - lo
        e = Engine.from_data(sub, shard_flags=L.SHARD_NO_PENALTY if r else 0)
        vr, gr_ = e.eval(par, 1)
        v, g = v + vr, g + gr_
        e.close()
    assert abs(v - ref_v) <= 1e-11 * abs(ref_v)
    assert grad_err(g, ref_g) <= 1e-9
    e1.close()


def test_sde_host_layer_with_user_H_maps_log_sigma_obs_off():
    """SDE$new(..., other_data = list(H = ...)): log_sigma_obs is fixed (R/sde.R:593-595) and the
    joint objective equals the oracle's on the same data list."""
    from smoothsde_b200.sde import SDE
    from smoothsde_b200.adfun import ADFun
    rng = np.random.default_rng(4)
    n = 400
    t = np.cumsum(rng.uniform(0.2, 2.0, n))
    x = np.cumsum(rng.normal(size=(n, 2)), axis=0)
    data = {"ID": np.repeat([1, 2], n // 2), "time": t, "x": x[:, 0], "y": x[:, 1]}
    Hs = np.stack([np.array([[0.05, 0.01], [0.01, 0.03]]) * (1 + 0.5 * rng.uniform()) for _ in range(n)], axis=2)
    sde = SDE(formulas={"mu1": "~1", "mu2": "~1", "tau": "~s(time, k=5, bs='cs')", "nu": "~1"}, data=data, type="CTCRW",
              response=["x", "y"], fixpar=["mu1", "mu2"], other_data={"H": Hs})
    tmb_dat, tmb_par, map_, random = sde.tmb_lists()
    assert map_["log_sigma_obs"] == [None]
    obj = ADFun(tmb_dat, tmb_par, map=map_, random=None)
    par = obj.par + 0.05 * rng.normal(size=obj.par.size)
    full = obj.full_from(par)
    ref = O.nllk(tmb_dat, full)
    assert abs(obj.fn(par) - ref) <= NLLK_RTOL * abs(ref)
    g_ref = obj.reduce_grad(oracle(tmb_dat, full)[1], obj._active)
    assert grad_err(obj.gr(par), g_ref) <= GRAD_RTOL
    obj.close()


def test_coupled_laplace_gradient_matches_differences_of_its_value():
    """The Laplace marginal over coeff_re (inner Newton with exact H_bb from tangent passes of the
    coupled kernels, third-derivative differences for its gradient) on a model with a user H_array:
    log_sigma_obs is mapped off as the reference does (R/sde.R:593-595)."""
    from smoothsde_b200.adfun import ADFun
    from test_gpu_laplace import split
    from smoothsde_b200 import synth
    dat, par, info = synth.make_problem("CTCRW", 3, 120, n_dim=2, seed=78, k=5, missing_frac=0.1)
    rng = np.random.default_rng(5)
    A = rng.normal(size=(info["n"], 2, 2)) * 0.2
    dat["H_array"] = np.ascontiguousarray((A @ A.transpose(0, 2, 1) + 0.02 * np.eye(2)).transpose(1, 2, 0))
    obj = ADFun(dat, split(dat, par, info), map={"coeff_fe": [None, None, 2, 3], "log_sigma_obs": [None]}, random="coeff_re")
    x = obj.par + 0.03 * np.arange(obj.par.size)
    f, g = obj._laplace.fn_gr(x)
    fd = np.empty(x.size)
    for j in range(x.size):
        e = np.zeros(x.size); e[j] = 1e-4
        fd[j] = (obj.fn(x + e) - obj.fn(x - e)) / 2e-4
    assert np.max(np.abs(g - fd)) <= 1e-6 * max(1.0, np.max(np.abs(fd))), (g, fd)
    obj.close()
