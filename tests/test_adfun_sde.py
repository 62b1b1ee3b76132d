"""CPU tests of the host mirror of the reference interface (smoothsde_b200/adfun.py, sde.py):
constructor validation as in tests/testthat/test_sde.R, TMB `map` semantics, the two objects of
SDE$setup(), fit() / logLik() plumbing.  The evaluator is the oracle-backed fake engine."""
import warnings

import numpy as np
import pytest

from fake_engine import OracleEngine, oracle_adfun
from smoothsde_b200 import simulate
from smoothsde_b200.adfun import ADFun
from smoothsde_b200.sde import SDE


def bm_data(n=200, seed=1, with_id=True):
    rng = np.random.default_rng(seed)
    t = np.arange(n, dtype=float)[None, :]
    x1 = np.cumsum(rng.normal(0, 0.1, n))
    z = simulate.simulate_bm(t, 0.3 * np.ones((1, n)), 0.8 * np.ones((1, n)), rng)[0]
    d = {"time": t[0], "Z": z, "x1": x1, "x2": rng.normal(size=n)}
    if with_id:
        d["ID"] = np.ones(n, dtype=int)
    return d


# ---- tests/testthat/test_sde.R:4-15
def test_constructor_works_for_bm_with_linear_covariate():
    sde = SDE(formulas={"mu": "~ x1", "sigma": "~ 1"}, data=bm_data(), type="BM", response="Z",
              adfun_factory=oracle_adfun)
    assert sde.coeff_fe().size == 3 and sde.coeff_re().size == 0


# ---- tests/testthat/test_sde.R:17-51
def test_missing_columns():
    d = bm_data(with_id=False)
    with pytest.warns(UserWarning, match="No ID column"):
        SDE(formulas={"mu": "~ 1", "sigma": "~ 1"}, data=d, type="BM", response="Z")
    d = bm_data()
    with pytest.raises(ValueError, match="'response' not found"):
        SDE(formulas={"mu": "~ 1", "sigma": "~ 1"}, data=d, type="BM", response="Y")
    with pytest.raises(KeyError):
        SDE(formulas={"mu": "~ x9", "sigma": "~ 1"}, data=d, type="BM", response="Z")
    d2 = {k: v for k, v in d.items() if k != "time"}
    with pytest.raises(ValueError, match="time column"):
        SDE(formulas={"mu": "~ 1", "sigma": "~ 1"}, data=d2, type="BM", response="Z")


# ---- tests/testthat/test_sde.R:53-72
def test_coefficient_bookkeeping():
    rng = np.random.default_rng(0)
    n = 100
    d = {"ID": np.repeat(np.arange(10), 10), "time": np.tile(np.arange(10.0), 10), "Z": rng.normal(size=n),
         "x1": rng.normal(size=n), "x2": rng.normal(size=n)}
    sde = SDE(formulas={"mu": "~ s(x1, k = 5, bs = 'ts') + x2", "sigma": "~ s(ID, bs = 're') + s(x2, k = 5, bs = 'ts')"},
              data=d, type="BM", response="Z")
    assert sde.coeff_fe().size == 3
    assert sde.coeff_re().size == 18
    assert sde.lambda_().size == 3
    assert sde.sdev().size == 3


def test_formula_and_type_validation():
    d = bm_data()
    with pytest.raises(ValueError, match="should be a list of length 2"):
        SDE(formulas={"mu": "~ 1"}, data=d, type="BM", response="Z")
    with pytest.raises(ValueError, match="components mu, sigma"):
        SDE(formulas={"sigma": "~ 1", "mu": "~ 1"}, data=d, type="BM", response="Z")
    with pytest.raises(ValueError, match="Unknown SDE type"):
        SDE(formulas=None, data=d, type="XYZ", response="Z")
    with pytest.raises(NotImplementedError):
        SDE(formulas=None, data=d, type="CIR", response="Z")
    with pytest.raises(ValueError, match="~1 for fixed"):
        SDE(formulas={"mu": "~ x1", "sigma": "~ 1"}, data=d, type="BM", response="Z", fixpar=["mu"])
    with pytest.raises(ValueError, match="'par0' should be of length 2"):
        SDE(formulas=None, data=d, type="BM", response="Z", par0=[1.0])
    sde = SDE(formulas=None, data=d, type="BM", response="Z", par0=[0.5, 2.0])
    assert np.allclose(sde.coeff_fe(), [0.5, np.log(2.0)])          # link applied, R/sde.R:156-159


def test_setup_lists_without_smooths_follow_the_reference():
    """R/sde.R:511-518: dummy S, ncol_re = 0, zero X_re column, coeff_re / log_lambda mapped off."""
    sde = SDE(formulas=None, data=bm_data(), type="BM", response="Z", adfun_factory=oracle_adfun)
    dat, par, map_, random = sde.tmb_lists()
    assert random is None and dat["S"].shape == (1, 1) and list(dat["ncol_re"]) == [0]
    assert dat["X_re"].shape[1] == 1 and dat["X_re"].nnz == 0
    assert map_["coeff_re"] == [None] and map_["log_lambda"] == [None]
    sde.setup()
    assert sde.tmb_obj().par.size == 2                               # only coeff_fe is free


def test_ctcrw_lists_and_fixpar_map():
    rng = np.random.default_rng(3)
    T, m = 2, 60
    t = simulate.make_times(T, m, rng, irregular=True)
    z = np.stack([simulate.simulate_ctcrw(t, np.zeros_like(t), np.ones_like(t), np.ones_like(t), rng) for _ in range(2)], -1)
    d = {"ID": np.repeat([7, 3], m), "time": t.ravel(), "x": z[..., 0].ravel(), "y": z[..., 1].ravel()}
    sde = SDE(formulas={"mu1": "~ 1", "mu2": "~ 1", "tau": "~ s(time, k = 5)", "nu": "~ 1"}, data=d, type="CTCRW",
              response=["x", "y"], par0=[0, 0, 1.5, 0.7], fixpar=["mu1", "mu2"], adfun_factory=oracle_adfun)
    dat, par, map_, random = sde.tmb_lists()
    assert list(par.keys()) == ["log_sigma_obs", "coeff_fe", "log_lambda", "coeff_re"]      # R/sde.R:590
    assert random == "coeff_re"
    assert dat["a0"].shape == (2, 4) and np.allclose(dat["a0"][:, [1, 3]], 0)                 # R/sde.R:574-580
    assert np.allclose(dat["a0"][1, [0, 2]], z[1, 0])
    assert np.allclose(np.diag(dat["P0"]), [1, 10, 1, 10])                                    # R/sde.R:584
    assert map_["coeff_fe"] == [None, None, 2, 3]                                             # R/sde.R:621-632
    # joint object: fixed entries stay at their initial value, gradient has the free entries only
    obj = oracle_adfun(dat, par, map=map_, random=None)
    assert obj.par.size == 1 + 2 + 1 + 4
    assert list(obj.names) == ["log_sigma_obs", "coeff_fe", "coeff_fe", "log_lambda"] + ["coeff_re"] * 4
    x = obj.par + 0.01
    full = obj.full_from(x)
    assert np.all(full[1:3] == 0.0)
    g = obj.gr(x)
    eng = OracleEngine(dat)
    _, g_full = eng.eval(full, 1)
    assert np.allclose(g, np.delete(g_full, [1, 2]))
    rep = obj.report()
    assert rep["aest_all"].shape == (2 * m, 4)


def test_map_ties_entries_together():
    dat, par, map_, _ = SDE(formulas={"mu1": "~ 1", "mu2": "~ 1", "sigma": "~ 1"}, data=dict(bm_data(), Z2=bm_data(seed=2)["Z"]),
                            type="BM", response=["Z", "Z2"]).tmb_lists()
    map_["coeff_fe"] = ["a", "a", "b"]                   # mu1 and mu2 share one coefficient
    obj = oracle_adfun(dat, par, map=map_)
    assert obj.par.size == 2
    x = np.array([0.2, -0.1])
    full = obj.full_from(x)
    assert full[0] == full[1] == 0.2
    _, g_full = OracleEngine(dat).eval(full, 1)
    assert np.allclose(obj.gr(x), [g_full[0] + g_full[1], g_full[2]])


def test_fit_bm_without_smooths_recovers_parameters_and_loglik():
    d = bm_data(n=400, seed=5)
    sde = SDE(formulas=None, data=d, type="BM", response="Z", par0=[0.0, 1.0], adfun_factory=oracle_adfun)
    res = sde.fit()
    assert res.success or res.status == 2
    mu_hat, sigma_hat = sde.coeff_fe()[0], np.exp(sde.coeff_fe()[1])
    dz = np.diff(d["Z"])
    assert abs(mu_hat - dz.mean()) < 1e-5                 # closed-form MLE for regular dt = 1
    assert abs(sigma_hat - dz.std()) < 1e-5
    from scipy.stats import norm
    assert abs(sde.logLik() - norm.logpdf(dz, mu_hat, sigma_hat).sum()) < 1e-6


def test_laplace_marginal_matches_dense_gaussian_integral():
    """For BM with a smooth on mu only, the joint nllk is quadratic in b, so the Laplace
    approximation is exact: f(theta) = -log int exp(-g(theta, b)) db."""
    rng = np.random.default_rng(2)
    n = 150
    d = bm_data(n=n, seed=9)
    sde = SDE(formulas={"mu": "~ s(time, k = 5)", "sigma": "~ 1"}, data=d, type="BM", response="Z",
              adfun_factory=oracle_adfun)
    dat, par, map_, random = sde.tmb_lists()
    obj = oracle_adfun(dat, par, map=map_, random=random)
    x = obj.par + np.array([0.1, -0.2, 0.3])              # mu intercept, log sigma, log lambda
    f = obj.fn(x)
    # brute force: g(theta, b) = c + q'b + b'Ab/2  ->  integral in closed form
    eng = OracleEngine(dat)
    nb = 4
    def g_of(b):
        return eng.eval(obj.full_from(x, b), 0)[0]
    g0 = g_of(np.zeros(nb))
    q = np.array([(g_of(e) - g_of(-e)) / 2 for e in np.eye(nb)])
    A = np.empty((nb, nb))
    for i in range(nb):
        for j in range(nb):
            ei, ej = np.eye(nb)[i], np.eye(nb)[j]
            A[i, j] = (g_of(ei + ej) - g_of(ei - ej) - g_of(-ei + ej) + g_of(-ei - ej)) / 4
    bhat = -np.linalg.solve(A, q)
    exact = g0 + q @ bhat + 0.5 * bhat @ A @ bhat + 0.5 * np.linalg.slogdet(A)[1] - 0.5 * nb * np.log(2 * np.pi)
    assert abs(f - exact) < 1e-6 * max(1.0, abs(exact))
    assert np.allclose(obj._laplace.b, bhat, atol=1e-6)


def test_laplace_gradient_matches_differences_of_the_marginal():
    """grad f = g_theta + 1/2 (w_theta - H_theta,b H_bb^-1 w_b), w from third-derivative differences
    of Hessian-vector products (laplace.py / csrc/ssde_laplace.cu), against central differences of
    the Laplace value itself.  OU: the joint nllk is NOT quadratic in b (tau smooth)."""
    from smoothsde_b200 import synth
    dat, par, info = synth.make_problem("OU", 2, 60, n_dim=1, seed=12, k=5, re_id=False)
    p_fe, n_s = info["p_fe"], info["n_s"]
    pars = {"coeff_fe": par[:p_fe], "log_lambda": par[p_fe:p_fe + n_s], "coeff_re": par[p_fe + n_s:]}
    obj = oracle_adfun(dat, pars, random="coeff_re")
    x = obj.par + 0.05 * np.arange(obj.par.size)
    f, g = obj._laplace.fn_gr(x)
    assert obj._laplace.last["converged"] == 1
    fd = np.empty(x.size)
    for j in range(x.size):
        e = np.zeros(x.size); e[j] = 1e-4
        fd[j] = (obj.fn(x + e) - obj.fn(x - e)) / 2e-4
    assert np.max(np.abs(g - fd)) <= 2e-5 * max(1.0, np.max(np.abs(fd))), (g, fd)


def test_he_and_sdreport_follow_tmb():
    """obj$he on the joint object honours `map`; sdreport's joint precision has TMB's block form and
    its fixed-effect block inverts to the marginal covariance (Schur complement identity)."""
    from smoothsde_b200 import synth
    dat, par, info = synth.make_problem("OU", 2, 60, n_dim=1, seed=12, k=5, re_id=False)
    p_fe, n_s = info["p_fe"], info["n_s"]
    pars = {"coeff_fe": par[:p_fe], "log_lambda": par[p_fe:p_fe + n_s], "coeff_re": par[p_fe + n_s:]}
    joint = oracle_adfun(dat, pars, map={"coeff_fe": [0, 1, None]})
    H = joint.he(joint.par)
    full = OracleEngine(dat).hessian(par)[2]
    keep = np.delete(np.arange(par.size), 2)
    assert np.allclose(H, full[np.ix_(keep, keep)], rtol=1e-9, atol=1e-9)
    obj = oracle_adfun(dat, pars, random="coeff_re")
    from scipy.optimize import minimize
    res = minimize(obj.fn, obj.par, jac=obj.gr, method="BFGS", options={"gtol": 1e-6})
    sd = obj.sdreport(res.x)
    nt, nb = res.x.size, sd["par_random"].size
    Q = sd["jointPrecision"]
    assert Q.shape == (nt + nb, nt + nb) and np.allclose(Q, Q.T)
    assert sd["names"] == ["coeff_fe"] * p_fe + ["log_lambda"] * n_s + ["coeff_re"] * nb
    cov = np.linalg.inv(Q)
    assert np.allclose(cov[:nt, :nt], sd["cov_fixed"], rtol=1e-6, atol=1e-10)       # marginal covariance of the fixed effects
    assert np.all(np.linalg.eigvalsh(Q) > 0)


def test_laplace_warm_start_chord_iterations_give_the_cold_start_result():
    from smoothsde_b200 import synth
    """The chord iterations that precondition with the factor of the previous mode must land on the same
    mode (and the same marginal value / gradient) as a cold start with full Newton steps."""
    from smoothsde_b200.laplace import LoopLaplace
    dat, par, info = synth.make_problem("CTCRW", 2, 60, n_dim=1, seed=3, k=5, missing_frac=0.1)
    warm = LoopLaplace(OracleEngine(dat))
    f0, g0, p0 = warm.eval(par, order=1)
    par2 = p0.copy()
    par2[0] += 0.05                       # log_sigma_obs moves, coeff_re starts at the previous mode
    par2[3] -= 0.03
    f_w, g_w, p_w = warm.eval(par2, order=1)
    assert warm.info["n_hess"] == 1       # the warm evaluation needed no Newton rebuild of H_bb
    f_c, g_c, p_c = LoopLaplace(OracleEngine(dat)).eval(par2, order=1)
    assert abs(f_w - f_c) <= 1e-9 * max(1.0, abs(f_c))
    assert np.max(np.abs(p_w - p_c)) <= 1e-7
    assert np.max(np.abs(g_w - g_c)) <= 1e-5 * max(1.0, np.max(np.abs(g_c)))
