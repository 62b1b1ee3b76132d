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
    ("OU_SSM", 3, 90, 0.1, 2),       # one-state-per-dimension Kalman models (nllk_ou_ssm.hpp, nllk_bm_ssm.hpp)
    ("OU_SSM", 1, 1500, 0.05, 1),    # several scan tiles
    ("BM_SSM", 4, 70, 0.2, 2),
    ("BM_SSM", 2, 1300, 0.0, 3),
]


@pytest.mark.parametrize("model,T,m,miss,nd", CASES)
def test_nllk_and_gradient_match_oracle(model, T, m, miss, nd):
    dat, par, info = synth.make_problem(model, T, m, missing_frac=miss, n_dim=nd, seed=7 + T + m)
    if model in ("CTCRW", "OU_SSM", "BM_SSM"):
        par = par.copy()
        par[1:1 + nd] = [0.3, -0.2, 0.1][:nd]     # exercise B*mu
    ref = O.nllk(dat, par)
    eng = Engine.from_data(dat)
    v0, _ = eng.eval(par, order=0)
    v, g = eng.eval(par, order=1)
    # order 0 sums the likelihood terms in the forward kernel's re-run, order 1 in the adjoint kernel's forward
    # recomputation (the re-run is skipped): the same filter, two compilations of it -- equal to rounding
    assert abs(v0 - v) <= 1e-14 * abs(v), (v0, v)
    assert abs(v - ref) <= NLLK_RTOL * abs(ref), (v, ref)
    g_ref = O.grad_complex_step(dat, par)
    assert grad_err(g, g_ref) <= GRAD_RTOL, (g, g_ref)
    # evaluating twice gives the same value (epoch / ticket reset)
    v2, g2 = eng.eval(par, order=1)
    assert abs(v2 - v) <= 1e-13 * abs(v)
    assert grad_err(g2, g) <= 1e-12
    eng.close()


@pytest.mark.parametrize("model,nd", [("OU_SSM", 2), ("BM_SSM", 1)])
def test_report_aest_ssm_matches_oracle(model, nd):
    dat, par, info = synth.make_problem(model, 3, 60, missing_frac=0.1, n_dim=nd, seed=8)
    par = par.copy()
    par[1:1 + nd] = [0.3, -0.2][:nd]
    p = O.split_par(dat, par)
    _, aest_ref = O._nllk_ssm(dat, **p, model=model, return_aest=True)
    eng = Engine.from_data(dat)
    eng.eval(par, order=0)
    aest = eng.report(dat["obs"].shape[0], nd, nd)
    ID = dat["ID"]
    last = np.r_[ID[1:] != ID[:-1], True]
    assert np.max(np.abs(aest[~last] - aest_ref[~last])) < 1e-9
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


@pytest.mark.parametrize("name", __import__("golden_util").names())
def test_engine_reproduces_golden_fixtures(name):
    import golden_util as G
    dat, gold = G.load(name)
    eng = Engine.from_data(dat)
    v, g = eng.eval(gold["par"], order=1)
    assert abs(v - gold["nllk"]) <= NLLK_RTOL * abs(gold["nllk"]), (v, gold["nllk"])
    assert grad_err(g, gold["grad"]) <= GRAD_RTOL
    eng.close()


EDGE_CASES = [
    ("CTCRW", 40, 3, 0.0, 2),        # many tiny tracks: several track starts inside one thread chunk
    ("CTCRW", 1, 2, 0.0, 1),         # a single transition-free track (two rows)
    ("CTCRW", 7, 301, 0.5, 2),       # half of the rows missing
    ("CTCRW", 3, 1024, 0.02, 2),     # track boundaries exactly on tile boundaries
    ("CTCRW", 1, 5000, 0.01, 2),     # one long track across several tiles of both kernels
    ("OU", 30, 4, 0.2, 2),
    ("BM", 1, 2, 0.0, 1),
    ("BM", 2, 1300, 0.05, 3),
]


@pytest.mark.parametrize("model,T,m,miss,nd", EDGE_CASES)
def test_edge_shapes_match_c_oracle(model, T, m, miss, nd):
    from oracle import oracle_c
    dat, par, info = synth.make_problem(model, T, m, missing_frac=miss, n_dim=nd, seed=1000 + T + m, k=5 if m < 20 else 10)
    if model == "CTCRW":
        par = par.copy()
        par[1:1 + nd] = [0.3, -0.2][:nd]
    ref_v, ref_g = oracle_c.COracle(dat).eval(par, True)
    eng = Engine.from_data(dat)
    v, g = eng.eval(par, order=1)
    assert abs(v - ref_v) <= NLLK_RTOL * max(abs(ref_v), 1.0), (v, ref_v)
    assert grad_err(g, ref_g) <= GRAD_RTOL
    eng.close()


def test_ragged_track_lengths_and_irregular_design():
    """Tracks of different lengths and a design whose rows do not share their columns (the
    per-nonzero column path of the layout)."""
    import scipy.sparse as sp
    from oracle import oracle_c
    rng = np.random.default_rng(5)
    lens = [3, 700, 41, 1, 260, 2]
    dat, par, info = synth.make_problem("CTCRW", 1, sum(lens), missing_frac=0.1, n_dim=2, seed=31)
    n = sum(lens)
    ID = np.repeat(np.arange(len(lens)), lens).astype(float)
    i0 = np.r_[0, np.cumsum(lens)[:-1]]
    obs = dat["obs"].copy()
    obs[i0] = np.nan_to_num(obs[i0])
    a0 = np.zeros((len(lens), 4))
    a0[:, 0], a0[:, 2] = obs[i0, 0], obs[i0, 1]
    p_re = 40
    Xre = sp.random(4 * n, p_re, density=0.04, random_state=9, format="csr") * 0.3
    Xre = sp.vstack([sp.csr_matrix((2 * n, p_re)),
                     sp.hstack([Xre[2 * n:3 * n, :20], sp.csr_matrix((n, 20))]),
                     sp.hstack([sp.csr_matrix((n, 20)), Xre[3 * n:, 20:]])], format="csr")
    dat = dict(dat, ID=ID, obs=obs, a0=a0, X_re=Xre, S=sp.identity(p_re, format="csr") * 2.0,
               ncol_re=np.array([20, 20]))
    par = np.r_[np.log(0.1), [0.1, -0.1, 0.0, 0.0], [0.3, -0.4], 0.2 * rng.standard_normal(p_re)]
    ref_v, ref_g = oracle_c.COracle(dat).eval(par, True)
    eng = Engine.from_data(dat)
    v, g = eng.eval(par, order=1)
    assert abs(v - ref_v) <= NLLK_RTOL * abs(ref_v), (v, ref_v)
    assert grad_err(g, ref_g) <= GRAD_RTOL
    eng.close()


def test_device_built_problem_matches_host_built_problem():
    """smoothsde_b200.devgen builds the packed layout with torch on the GPU (used by bench.py at
    1e8 rows); the same rows passed through ssde_create must give the same objective."""
    import scipy.sparse as sp
    import torch
    from smoothsde_b200 import design as D
    from smoothsde_b200 import devgen
    eng, par, info = devgen.make_ctcrw_device(5, 700, seed=3, device=0)
    v, g = eng.eval(par, order=1)
    n, k = info["n"], 10
    t = info["tensors"]
    times = t["times"].cpu().numpy()
    lo, hi = info["knots"]
    Bz = D.bspline_basis(times, k, lo, hi) @ info["Zc"]
    one = sp.csr_matrix(np.ones((n, 1)))
    X_fe = sp.block_diag([one, one, one, one], format="csr")
    X_re = sp.block_diag([sp.csr_matrix((n, 0)), sp.csr_matrix((n, 0)), sp.csr_matrix(Bz), sp.csr_matrix(Bz)], format="csr")
    dat = {"type": "CTCRW", "ID": np.repeat(np.arange(5), 700).astype(float), "times": times,
           "obs": t["obs"].cpu().numpy(), "X_fe": X_fe, "X_re": X_re, "S": info["S"],
           "ncol_re": np.array([9, 9]), "include_penalty": 1, "a0": info["a0"],
           "P0": np.diag([1.0, 10.0, 1.0, 10.0])}
    eng2 = Engine.from_data(dat)
    v2, g2 = eng2.eval(par, order=1)
    assert abs(v - v2) <= 1e-12 * abs(v2), (v, v2)
    assert grad_err(g, g2) <= 1e-10
    eng.close(); eng2.close()
    del t
    torch.cuda.empty_cache()


# ---------------------------------------------------------------------------------------------
# second order: exact Hessian-vector products (tangent pass) and the joint Hessian (obj$he)
# ---------------------------------------------------------------------------------------------
HESS_RTOL = 1e-6


def hess_err(H, H_ref):
    scale = np.maximum(np.abs(H_ref), 1e-3 * np.max(np.abs(H_ref)))
    return np.max(np.abs(H - H_ref) / scale)


@pytest.mark.parametrize("name", __import__("golden_util").names())
def test_joint_hessian_reproduces_golden_fixtures(name):
    import golden_util as G
    dat, gold = G.load(name)
    eng = Engine.from_data(dat)
    v, g, H = eng.hessian(gold["par"])
    assert abs(v - gold["nllk"]) <= NLLK_RTOL * abs(gold["nllk"])
    assert grad_err(g, gold["grad"]) <= GRAD_RTOL
    assert hess_err(H, gold["hess"]) <= HESS_RTOL, np.max(np.abs(H - gold["hess"]))
    # raw columns (independent tangent passes) are symmetric to rounding
    _, _, Hraw = eng.hvp(gold["par"], np.eye(gold["par"].size))
    assert np.max(np.abs(Hraw - Hraw.T)) <= 1e-9 * np.max(np.abs(Hraw))
    # the plain evaluation is unaffected by tangent passes in between
    v1, g1 = eng.eval(gold["par"], order=1)
    assert abs(v1 - v) <= 1e-13 * abs(v) and grad_err(g1, g) <= 1e-12
    eng.close()


@pytest.mark.parametrize("model,T,m,miss,nd", [
    ("CTCRW", 2, 1500, 0.05, 2),     # several scan tiles per kernel
    ("CTCRW", 40, 3, 0.0, 2),        # many tiny tracks
    ("CTCRW", 5, 41, 0.3, 1),
    ("OU", 30, 40, 0.2, 2),
    ("BM", 2, 1300, 0.05, 3),
])
def test_hessian_vector_products_match_oracle_differences(model, T, m, miss, nd):
    from oracle import oracle_c
    dat, par, info = synth.make_problem(model, T, m, missing_frac=miss, n_dim=nd, seed=77 + T + m, k=5 if m < 20 else 10)
    if model == "CTCRW":
        par = par.copy()
        par[1:1 + nd] = [0.3, -0.2][:nd]
    rng = np.random.default_rng(2)
    dirs = rng.normal(size=(par.size, 3))
    dirs[:, 2] = 0.0
    dirs[-1, 2] = 1.0                                # a pure coeff_re unit direction
    co = oracle_c.COracle(dat)
    eng = Engine.from_data(dat)
    v, g, hv = eng.hvp(par, dirs)
    ref_v, ref_g = co.eval(par, True)
    assert abs(v - ref_v) <= NLLK_RTOL * max(abs(ref_v), 1.0)
    assert grad_err(g, ref_g) <= GRAD_RTOL
    k = 1e-3
    for c in range(dirs.shape[1]):
        d = dirs[:, c]
        d1 = (co.eval(par + k * d)[1] - co.eval(par - k * d)[1]) / (2 * k)
        d2 = (co.eval(par + 0.5 * k * d)[1] - co.eval(par - 0.5 * k * d)[1]) / k
        ref = (4 * d2 - d1) / 3
        assert hess_err(hv[:, c], ref) <= HESS_RTOL, (c, np.max(np.abs(hv[:, c] - ref)))
    eng.close()


# ---------------------------------------------------------------------------------------------
# boundary behaviour (include/smoothsde_b200.h)
# ---------------------------------------------------------------------------------------------
def test_partially_missing_row_is_rejected_for_kalman_models():
    """obs[i, 0] observed, obs[i, 1] NA: the reference's objective is NaN for such a row (only column
    0 is tested, nllk_ctcrw.hpp:214; tests/test_ref.py::test_na_semantics_of_the_reference) -- the
    engine refuses the data instead of silently filtering y = 0."""
    from smoothsde_b200 import _lib
    dat, par, info = synth.make_problem("CTCRW", 2, 30, missing_frac=0.0, n_dim=2, seed=5)
    obs = dat["obs"].copy()
    obs[7, 1] = np.nan
    with pytest.raises(_lib.EngineError) as ei:
        Engine.from_data(dict(dat, obs=obs))
    assert ei.value.code == 2 and "another column is NA" in str(ei.value)
    # column 0 NA, column 1 observed: the row simply counts as missing, as in the reference
    obs = dat["obs"].copy()
    obs[7, 0] = np.nan
    eng = Engine.from_data(dict(dat, obs=obs))
    v, _ = eng.eval(par, order=1)
    ref = O.nllk(dict(dat, obs=obs), par)
    assert abs(v - ref) <= NLLK_RTOL * abs(ref)
    eng.close()


def test_nonpositive_innovation_variance_sets_the_numeric_status():
    """SSDE_ERR_NUMERIC 'F <= 0 in the filter': P0 = 0 and sigma_obs^2 underflowing to 0 give F = 0 on
    the first filtered row; the reference would take its detF <= 0 branch (nllk_ctcrw.hpp:226-228),
    which the engine does not build -- it must say so instead of returning a number."""
    from smoothsde_b200 import _lib
    dat, par, info = synth.make_problem("CTCRW", 2, 30, missing_frac=0.0, n_dim=1, seed=6)
    dat = dict(dat, P0=np.zeros((2, 2)))
    par = par.copy()
    par[0] = -400.0                                   # sigma_obs^2 = exp(-800) = 0
    eng = Engine.from_data(dat)
    with pytest.raises(_lib.EngineError) as ei:
        eng.eval(par, order=1)
    assert ei.value.code == 5 and "F <= 0" in str(ei.value)
    par[0] = np.log(0.1)
    v, _ = eng.eval(par, order=1)                     # the handle stays usable
    assert np.isfinite(v)
    eng.close()


# ---------------------------------------------------------------------------------------------
# one-pass data-term Hessian of the BM / OU models: X' W X + lambda S (ssde_hess_theta_device)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("model,T,m,nd,re_id,k", [("OU", 6, 300, 1, True, 10), ("OU", 3, 500, 2, False, 8), ("BM", 4, 400, 2, False, 10),
                                                  ("BM", 2, 700, 3, False, 6), ("OU", 10, 120, 1, True, 5)])
def test_one_pass_hessian_equals_tangent_pass_hessian(model, T, m, nd, re_id, k):
    """H = X' W X with the exact per-row second-derivative blocks (DualN through sde_row) against the joint
    Hessian assembled column by column from tangent passes (itself checked against the reference's
    reverse-over-forward Hessian in tests/test_gpu_ref.py)."""
    dat, par, info = synth.make_problem(model, T, m, missing_frac=0.1, n_dim=nd, seed=90 + T + m, re_id=re_id, k=k)
    eng = Engine.from_data(dat)
    _, _, H = eng.hessian(par)
    Ht = eng.hess_theta(par)
    o_fe, p_fe = eng.layout["coeff_fe"]
    o_re, p_re = eng.layout["coeff_re"]
    idx = np.r_[o_fe + np.arange(p_fe), o_re + np.arange(p_re)]
    ref = H[np.ix_(idx, idx)]
    assert Ht.shape == ref.shape
    assert np.max(np.abs(Ht - Ht.T)) <= 1e-12 * np.max(np.abs(ref))
    assert np.max(np.abs(Ht - ref)) <= 1e-10 * np.max(np.abs(ref)), np.max(np.abs(Ht - ref)) / np.max(np.abs(ref))
    eng.close()


def test_one_pass_hessian_on_the_device_built_ou_shape():
    """Random intercepts per track, warp-tiles that straddle two tracks (25 slots), p_theta > the hot set."""
    from smoothsde_b200 import devgen
    eng, par, info = devgen.make_ou_device(12, 300, device=0)
    rng = np.random.default_rng(5)
    par = par.copy()
    par[:3] = [0.2, 0.1, np.log(1.3)]
    par[3:7] = [0.3, -0.4, 0.2, 0.5]
    par[7:] = 0.2 * rng.standard_normal(par.size - 7)
    _, _, H = eng.hessian(par)
    Ht = eng.hess_theta(par)
    o_fe, p_fe = eng.layout["coeff_fe"]
    o_re, p_re = eng.layout["coeff_re"]
    idx = np.r_[o_fe + np.arange(p_fe), o_re + np.arange(p_re)]
    ref = H[np.ix_(idx, idx)]
    assert np.max(np.abs(Ht - ref)) <= 1e-10 * np.max(np.abs(ref)), np.max(np.abs(Ht - ref)) / np.max(np.abs(ref))
    with pytest.raises(Exception):
        dat, p2, _ = synth.make_problem("CTCRW", 2, 50, n_dim=1, seed=1)
        e2 = Engine.from_data(dat)
        try:
            e2.hess_theta(p2)                      # Kalman models: SSDE_ERR_UNSUPPORTED
        finally:
            e2.close()
    eng.close()
