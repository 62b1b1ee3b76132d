"""Seeded synthetic problems of the shapes named in BASELINE.json / SURVEY.md section 8(d).

Returns the same *data list* / *parameter list* that SDE$setup() hands to TMB::MakeADFun
(R/sde.R:491-670): ``type, ID, times, obs, X_fe, X_re, S, ncol_re, include_penalty`` plus, for
CTCRW, ``a0`` (first position of each track, zero velocity; R/sde.R:574-580) and
``P0 = diag(rep(c(1, 10), n_dim))`` (R/sde.R:584).
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np

from . import design as _design
from . import simulate as _sim


def true_pars(s):
    """Smooth 'true' parameters in normalised time s in [0, 1]."""
    tau = np.exp(0.5 * np.sin(2 * np.pi * s))
    nu = np.exp(0.3 * np.cos(2 * np.pi * s))
    return tau, nu


def make_problem(model, n_tracks, n_steps, seed=20260101, irregular=None, missing_frac=0.0,
                 k=10, re_id=None, sigma_obs=0.1, n_dim=None):
    """Build one synthetic problem.

    model: "BM" | "OU" | "CTCRW" | "BM_SSM" | "OU_SSM" (the last two: BM / OU paths observed with
    measurement error sigma_obs; a0 = first observation, P0 = diag(10), R/sde.R:547-557).
    Formulas follow SURVEY.md 8(d):
      BM     mu, sigma ~ s(time, k)                          (C1)
      OU     mu, tau ~ s(time, k) [+ s(ID, bs="re")], kappa ~ 1   (C2)
      CTCRW  mu1 = mu2 = 0 fixed, tau, nu ~ s(time, k)       (C3/C4)
    Returns (dat, par, info).
    """
    rng = np.random.default_rng(seed)
    ssm = model in ("BM_SSM", "OU_SSM")
    out_type = model
    if ssm:
        model = model[:2]
    if irregular is None:
        irregular = model == "CTCRW" or ssm
    if re_id is None:
        re_id = model == "OU" and n_tracks > 1 and not ssm
    if n_dim is None:
        n_dim = 2 if model == "CTCRW" else 1
    T, m = n_tracks, n_steps
    n = T * m
    times = _sim.make_times(T, m, rng, irregular)
    s = times / times[:, -1:].clip(min=1e-300)
    tau, nu = true_pars(s)
    ID = np.repeat(np.arange(1, T + 1), m)
    tflat = times.ravel()
    data = {"ID": ID, "time": tflat}

    sm = f"s(time, k = {k}, bs = 'cs')"
    if model == "BM":
        mu = 0.1 * np.sin(2 * np.pi * s)
        sig = nu
        z = np.stack([_sim.simulate_bm(times, mu, sig, rng) for _ in range(n_dim)], axis=-1)
        formulas = OrderedDict([(f"mu{i + 1}" if n_dim > 1 else "mu", "~ " + sm)
                                for i in range(n_dim)] + [("sigma", "~ " + sm)])
        beta0 = [0.0] * n_dim + [0.0]
    elif model == "OU":
        mu = 2.0 * np.sin(2 * np.pi * s)
        kappa = np.full_like(tau, 1.5)
        z = np.stack([_sim.simulate_ou(times, mu, tau, kappa, rng) for _ in range(n_dim)], axis=-1)
        rhs = "~ " + sm + (" + s(ID, bs = 're')" if re_id else "")
        formulas = OrderedDict([(f"mu{i + 1}" if n_dim > 1 else "mu", rhs) for i in range(n_dim)]
                               + [("tau", rhs), ("kappa", "~ 1")])
        beta0 = [0.0] * n_dim + [0.0, np.log(1.5)]
    elif model == "CTCRW":
        mu = np.zeros_like(tau)
        zs = [_sim.simulate_ctcrw(times, mu, tau, nu, rng) for _ in range(n_dim)]
        z = np.stack(zs, axis=-1)
        z = z + sigma_obs * rng.standard_normal(z.shape)
        formulas = OrderedDict([(f"mu{i + 1}", "~ 1") for i in range(n_dim)]
                               + [("tau", "~ " + sm), ("nu", "~ " + sm)])
        beta0 = [0.0] * n_dim + [0.0, 0.0]
    else:
        raise ValueError("Unknown SDE type")
    if ssm:
        z = z + sigma_obs * rng.standard_normal(z.shape)
    obs = z.reshape(n, n_dim).copy()
    if missing_frac > 0:
        miss = rng.random(n) < missing_frac
        miss[::m] = False            # keep the first observation of each track (a0 needs it)
        obs[miss, :] = np.nan

    des = _design.make_design(formulas, data, n)
    dat = {
        "type": out_type, "ID": ID.astype(float), "times": tflat, "obs": obs,
        "X_fe": des.X_fe, "X_re": des.X_re, "S": des.S, "ncol_re": des.ncol_re,
        "include_penalty": 1,
    }
    if model == "CTCRW":
        i0 = np.arange(0, n, m)
        a0 = np.zeros((T, 2 * n_dim))
        for d in range(n_dim):
            a0[:, 2 * d] = obs[i0, d]
        dat["a0"] = a0
        dat["P0"] = np.diag(np.tile([1.0, 10.0], n_dim))
    if ssm:
        dat["a0"] = obs[np.arange(0, n, m)].copy()
        dat["P0"] = 10.0 * np.eye(n_dim)

    # parameter vector for parity checks: beta at the link of the true means, b ~ N(0, 0.1^2),
    # log lambda = 0, log sigma_obs = log 0.1 (seed + 1)
    prng = np.random.default_rng(seed + 1)
    p_fe, p_re = des.X_fe.shape[1], des.X_re.shape[1]
    coeff_fe = np.asarray(beta0, dtype=float)
    assert coeff_fe.size == p_fe, (coeff_fe.size, p_fe)
    coeff_re = 0.1 * prng.standard_normal(p_re)
    log_lambda = np.zeros(des.ncol_re.size)
    pieces = ([np.array([np.log(sigma_obs)])] if (model == "CTCRW" or ssm) else []) + \
        [coeff_fe, log_lambda, coeff_re]
    par = np.concatenate(pieces)
    info = {"design": des, "formulas": formulas, "n": n, "n_dim": n_dim, "n_tracks": T,
            "n_steps": m, "p_fe": p_fe, "p_re": p_re, "n_s": des.ncol_re.size}
    return dat, par, info
