"""The oracle is pinned to the reference's OWN code.

oracle/_ref/libsmoothsde_ref.so is /root/reference/src/smoothSDE.cpp + src/nllk/*.hpp, unmodified,
compiled where they lie against the TMB stand-in oracle/tmb_shim/TMB.hpp (oracle/Makefile).  These
tests check, on every committed fixture and for all five built model types, that

    reference objective  ==  numpy restatement  ==  C restatement  ==  stored reference vectors

value 1e-13, gradient 1e-10 (the north-star tolerances are 1e-10 / 1e-7), Hessian 1e-9 of max|H|.
The stored vectors (`ref_*` in tests/golden/*.npz, made by tests/golden/make_ref_golden.py) are
what the -m gpu tests compare the CUDA engine with."""
import numpy as np
import pytest

import golden_util as G
from oracle import oracle_np as O
from oracle import oracle_ref
from smoothsde_b200 import synth

pytestmark = pytest.mark.skipif(not oracle_ref.available(), reason="oracle/_ref not built and /root/reference absent")


def gerr(g, g_ref):
    scale = np.maximum(np.abs(g_ref), 1e-3 * np.max(np.abs(g_ref)))
    return float(np.max(np.abs(g - g_ref) / scale))


@pytest.mark.parametrize("name", G.names())
def test_reference_equals_restatements_on_fixtures(name):
    dat, gold = G.load(name)
    assert "ref_nllk" in gold, "fixture lacks reference vectors: run tests/golden/make_ref_golden.py"
    R = oracle_ref.RefOracle(dat)
    v, g = R.eval(gold["par"])
    assert R.nllk(gold["par"]) == v                      # Type = double and Type = AD agree bit for bit
    # the library reproduces the committed reference vectors
    assert abs(v - gold["ref_nllk"]) <= 1e-14 * abs(v)
    assert gerr(g, gold["ref_grad"]) <= 1e-12
    # numpy restatement (value + complex-step gradient, stored in the fixture by make_golden.py)
    assert abs(O.nllk(dat, gold["par"]) - v) <= 1e-13 * abs(v)
    assert abs(gold["nllk"] - v) <= 1e-13 * abs(v)
    assert gerr(gold["grad"], g) <= 1e-10
    # independent known answer stored with the fixture (dense MVN / scipy logpdf sums)
    assert abs(gold["known_answer"] - v) <= 1e-9 * abs(v)
    if G.c_oracle_covers(dat) and dat["type"] in ("BM", "OU", "CTCRW"):
        from oracle import oracle_c
        cv, cg = oracle_c.COracle(dat).eval(gold["par"], True)
        assert abs(cv - v) <= 1e-13 * abs(v)
        assert gerr(cg, g) <= 1e-10


@pytest.mark.parametrize("name", [n for n in G.names() if n in ("ctcrw_d2_3x40", "ou_d1_4x60", "bm_ssm_d1_3x40", "ctcrw_d2_userH_2x40", "ou_d1_decay_3x50")])
def test_reference_hessian_matches_fixture_hessians(name):
    dat, gold = G.load(name)
    R = oracle_ref.RefOracle(dat)
    H = R.hessian(gold["par"])
    assert np.max(np.abs(H - gold["ref_hess"])) <= 1e-12 * np.abs(H).max()
    # the restatements' Hessian (Richardson FD of the analytic gradient / complex-step FD)
    assert np.max(np.abs(H - gold["hess"])) <= 1e-7 * np.abs(H).max()
    # Hessian-vector product along a random direction = H v
    rng = np.random.default_rng(1)
    d = rng.standard_normal(gold["par"].size)
    _, g, hv = R.hvp(gold["par"], d)
    assert gerr(g, gold["ref_grad"]) <= 1e-12
    assert np.max(np.abs(hv - H @ d)) <= 1e-10 * np.abs(H @ d).max()


@pytest.mark.parametrize("model,nd", [("CTCRW", 2), ("OU_SSM", 2), ("BM_SSM", 1)])
def test_reported_states_match(model, nd):
    dat, par, info = synth.make_problem(model, 3, 50, missing_frac=0.1, n_dim=nd, seed=31)
    par = par.copy()
    par[1:1 + nd] = [0.3, -0.2][:nd]
    p = O.split_par(dat, par)
    if model == "CTCRW":
        _, aest = O.nllk_ctcrw(dat, **p, return_aest=True)
    else:
        _, aest = O._nllk_ssm(dat, **p, model=model, return_aest=True)
    ref = oracle_ref.RefOracle(dat).aest(par)
    assert ref.shape == aest.shape
    assert np.max(np.abs(ref - aest)) <= 1e-11 * max(1.0, np.abs(aest).max())


def test_na_semantics_of_the_reference():
    """R_IsNA tests column 0 only (nllk_ctcrw.hpp:214): a row whose column 0 is observed and whose
    column 1 is NA poisons the objective with NaN -- the engine rejects such data at ssde_create."""
    dat, par, info = synth.make_problem("CTCRW", 2, 30, missing_frac=0.0, n_dim=2, seed=5)
    obs = dat["obs"].copy()
    obs[7, 1] = np.nan
    v = oracle_ref.RefOracle(dict(dat, obs=obs)).nllk(par)
    assert np.isnan(v)
    assert np.isnan(O.nllk(dict(dat, obs=obs), par))
    # column 0 NA and column 1 observed: the row counts as missing, finite objective
    obs = dat["obs"].copy()
    obs[7, 0] = np.nan
    v = oracle_ref.RefOracle(dict(dat, obs=obs)).nllk(par)
    assert np.isfinite(v) and abs(v - O.nllk(dict(dat, obs=obs), par)) <= 1e-13 * abs(v)


def test_no_smooth_model_and_include_penalty_flag():
    """No smooths: S = 0 (1 x 1), ncol_re = 0, X_re a zero column (R/sde.R:511-518); and
    include_penalty = 0 drops the penalty for nllk_sde but not for the Kalman models."""
    import scipy.sparse as sp
    dat, par, info = synth.make_problem("OU", 3, 60, missing_frac=0.1, n_dim=1, seed=9)
    v1 = oracle_ref.RefOracle(dat).nllk(par)
    v0 = oracle_ref.RefOracle(dict(dat, include_penalty=0)).nllk(par)
    assert abs(v0 - O.nllk(dict(dat, include_penalty=0), par)) <= 1e-13 * abs(v0)
    assert abs((v1 - v0) - O.penalty_sde(dat, O.split_par(dat, par)["log_lambda"], O.split_par(dat, par)["coeff_re"])) <= 1e-10
    n = dat["obs"].shape[0]
    nos = dict(dat, X_re=sp.csr_matrix((dat["X_fe"].shape[0], 1)), S=sp.csr_matrix((1, 1)), ncol_re=np.array([0]))
    p = O.split_par(dat, par)
    par_nos = np.concatenate([p["coeff_fe"], [0.0], [0.0]])
    v = oracle_ref.RefOracle(nos).nllk(par_nos)
    assert abs(v - O.nllk(nos, par_nos)) <= 1e-13 * abs(v)
    assert n > 0


def test_track_parallel_evaluation_equals_stacked():
    for model, nd in (("CTCRW", 2), ("OU", 1)):
        dat, par, info = synth.make_problem(model, 5, 40, missing_frac=0.1, n_dim=nd, seed=77)
        v, g = oracle_ref.RefOracle(dat).eval(par)
        P = oracle_ref.RefOracleParallel(dat, nthreads=3)
        pv, pg = P.eval(par)
        P.close()
        assert abs(pv - v) <= 1e-12 * abs(v)
        assert gerr(pg, g) <= 1e-10


def test_unknown_type_raises_like_the_reference():
    dat, par, info = synth.make_problem("BM", 1, 20, n_dim=1, seed=2)
    with pytest.raises(RuntimeError, match="Unknown SDE type"):       # smoothSDE.cpp:25
        oracle_ref.RefOracle(dict(dat, type="XYZ")).nllk(par)


def test_long_tracks_with_restarting_clocks():
    """Thousands of rows per track, several tracks whose clocks restart at 0: the reference's
    discarded cross-track prediction overflows (negative dt) and must not poison the reverse sweep;
    the 2-D filter's zero-valued cross-covariances are still AD variables."""
    from oracle import oracle_c
    dat, par, info = synth.make_problem("CTCRW", 3, 2500, missing_frac=0.05, n_dim=2, seed=3003)
    par = par.copy()
    par[1:3] = [0.3, -0.2]
    rv, rg = oracle_ref.RefOracle(dat).eval(par)
    cv, cg = oracle_c.COracle(dat).eval(par, True)
    assert np.isfinite(rg).all()
    assert abs(rv - cv) <= 1e-13 * abs(cv)
    assert gerr(rg, cg) <= 1e-10
