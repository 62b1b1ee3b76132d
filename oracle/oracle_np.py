"""numpy restatement of smoothSDE's penalised negative log-likelihood (TEST INFRASTRUCTURE ONLY).

This file is a CPU *oracle*: a line-by-line restatement, in numpy, of the objective body that the
reference builds with TMB.  It is imported only by tests/, by __graft_entry__.smoke() and by
bench.py's cpu_baseline leg.  The product path (smoothsde_b200/) never imports it.

PARITY UNPINNED (w.r.t. TMB): the reference ships no numeric test, golden vector or fixture for
this path (tests/testthat/test_sde.R:1-72 only checks vector lengths) and neither R nor TMB exist
in this image, so the oracle cannot be run against the real TMB objective.  It is pinned instead
by independent known-answer checks (dense multivariate-normal density, scipy.stats.norm.logpdf
sums, complex-step vs hand adjoint, see tests/test_oracle.py and tests/golden/).

Reference files followed (paths relative to /root/reference):
  * src/nllk/nllk_ctcrw.hpp:12-91   det / makeH / makeT / makeQ / makeB
  * src/nllk/nllk_ctcrw.hpp:102-283 nllk_ctcrw (Kalman loop + penalty)
  * src/nllk/nllk_sde.hpp:15-127    nllk_sde (BM / OU loop + penalty with constants)
  * src/nllk/tr_dens.hpp:18-75      tr_dens (BM :32-37, OU :45-52)
  * src/smoothSDE.cpp:9-28          dispatch on `type`

All functions are dtype-generic: pass complex parameters to get complex-step derivatives
(the objective is analytic in the parameters; branches use the real part).
"""
from __future__ import annotations

import math

import numpy as np

LOG_SQRT_2PI = 0.5 * math.log(2.0 * math.pi)
KALMAN_TYPES = ("CTCRW", "OU_SSM", "BM_SSM")


def _re(x):
    return x.real if np.iscomplexobj(x) else x


def _isna(x) -> bool:
    """R_IsNA stand-in: any NaN counts as missing (the C ABI documents the same rule)."""
    return bool(np.isnan(_re(x)))


# --------------------------------------------------------------------------------------------
# CTCRW helpers, nllk_ctcrw.hpp:12-91
# --------------------------------------------------------------------------------------------
def det_small(M):
    """nllk_ctcrw.hpp:12-24 -- closed form for n_dim<=2, exp(logdet) otherwise."""
    nd = M.shape[1]
    if nd == 1:
        return M[0, 0]
    if nd == 2:
        return M[0, 0] * M[1, 1] - M[1, 0] * M[0, 1]
    sign, ld = np.linalg.slogdet(M)
    return sign * np.exp(ld)


def makeH_ctcrw(sigma_obs, n_dim, dtype):
    """nllk_ctcrw.hpp:30-38"""
    H = np.zeros((n_dim, n_dim), dtype=dtype)
    for i in range(n_dim):
        H[i, i] = sigma_obs * sigma_obs
    return H


def makeT_ctcrw(beta, dt, n_dim, dtype):
    """nllk_ctcrw.hpp:45-55"""
    T = np.zeros((2 * n_dim, 2 * n_dim), dtype=dtype)
    for i in range(n_dim):
        T[2 * i, 2 * i] = 1
        T[2 * i, 2 * i + 1] = (1 - np.exp(-beta * dt)) / beta
        T[2 * i + 1, 2 * i + 1] = np.exp(-beta * dt)
    return T


def makeQ_ctcrw(beta, sigma, dt, n_dim, dtype):
    """nllk_ctcrw.hpp:63-75 (same cancellation-prone closed form as the reference)."""
    Q = np.zeros((2 * n_dim, 2 * n_dim), dtype=dtype)
    for i in range(n_dim):
        Q[2 * i, 2 * i] = (sigma / beta) * (sigma / beta) * (
            dt - 2 / beta * (1 - np.exp(-beta * dt)) + 1 / (2 * beta) * (1 - np.exp(-2 * beta * dt))
        )
        Q[2 * i, 2 * i + 1] = sigma * sigma / (2 * beta * beta) * (
            1 - 2 * np.exp(-beta * dt) + np.exp(-2 * beta * dt)
        )
        Q[2 * i + 1, 2 * i] = Q[2 * i, 2 * i + 1]
        Q[2 * i + 1, 2 * i + 1] = sigma * sigma / (2 * beta) * (1 - np.exp(-2 * beta * dt))
    return Q


def makeB_ctcrw(beta, dt, n_dim, dtype):
    """nllk_ctcrw.hpp:82-91"""
    B = np.zeros((2 * n_dim, n_dim), dtype=dtype)
    for i in range(n_dim):
        B[2 * i, i] = dt - (1 - np.exp(-beta * dt)) / beta
        B[2 * i + 1, i] = 1 - np.exp(-beta * dt)
    return B


def linear_predictor(dat, coeff_fe, coeff_re):
    """par_vec = X_fe*coeff_fe + X_re*coeff_re ; par_mat[i, j] = par_vec[j*n + i]
    (nllk_ctcrw.hpp:143-149, nllk_sde.hpp:61-67)."""
    n = dat["obs"].shape[0]
    par_vec = dat["X_fe"] @ np.asarray(coeff_fe) + dat["X_re"] @ np.asarray(coeff_re)
    par_vec = np.asarray(par_vec).ravel()
    n_par = par_vec.size // n
    return par_vec.reshape(n_par, n).T  # column j = segment j


def _block(S, a, b):
    blk = S[a:b, a:b]
    return np.asarray(blk.todense()) if hasattr(blk, "todense") else np.asarray(blk, dtype=float)


def penalty_kalman(dat, log_lambda, coeff_re):
    """nllk_ctcrw.hpp:254-280 -- no 2*pi / log-det constants, ignores include_penalty."""
    ncol_re = np.atleast_1d(np.asarray(dat["ncol_re"], dtype=np.int64))
    pen = 0.0
    if ncol_re[0] > 0:
        S = dat["S"]
        S_start = 0
        for i in range(ncol_re.size):
            Sn = int(ncol_re[i])
            this_S = _block(S, S_start, S_start + Sn)
            b = coeff_re[S_start:S_start + Sn]
            pen = pen - 0.5 * Sn * log_lambda[i] + 0.5 * np.exp(log_lambda[i]) * (b @ (this_S @ b))
            S_start += Sn
    return pen


def penalty_sde(dat, log_lambda, coeff_re):
    """nllk_sde.hpp:89-124 -- includes 0.5*Sn*log(2*pi) - 0.5*log det S_i; needs full-rank S_i."""
    ncol_re = np.atleast_1d(np.asarray(dat["ncol_re"], dtype=np.int64))
    pen = 0.0
    if ncol_re[0] > 0 and int(dat.get("include_penalty", 1)) != 0:
        S = dat["S"]
        S_start = 0
        for i in range(ncol_re.size):
            Sn = int(ncol_re[i])
            this_S = _block(S, S_start, S_start + Sn)
            b = coeff_re[S_start:S_start + Sn]
            sign, logdetS = np.linalg.slogdet(this_S)
            log_det = -logdetS  # log det(S^-1), nllk_sde.hpp:109-111
            pen = (pen + 0.5 * Sn * math.log(2 * math.pi) + 0.5 * log_det
                   - 0.5 * Sn * log_lambda[i]
                   + 0.5 * np.exp(log_lambda[i]) * (b @ (this_S @ b)))
            S_start += Sn
    return pen


# --------------------------------------------------------------------------------------------
# nllk_ctcrw, nllk_ctcrw.hpp:102-283
# --------------------------------------------------------------------------------------------
def nllk_ctcrw(dat, log_sigma_obs, coeff_fe, log_lambda, coeff_re, return_aest=False):
    ID = np.asarray(dat["ID"])
    times = np.asarray(dat["times"], dtype=float)
    obs = np.asarray(dat["obs"], dtype=float)
    a0 = np.asarray(dat["a0"], dtype=float)
    P0 = np.asarray(dat["P0"], dtype=float)
    H_array = dat.get("H_array", None)
    n, n_dim = obs.shape

    coeff_fe = np.asarray(coeff_fe)
    coeff_re = np.asarray(coeff_re)
    log_lambda = np.asarray(log_lambda)
    dtype = np.result_type(coeff_fe.dtype, coeff_re.dtype, np.asarray(log_sigma_obs).dtype,
                           log_lambda.dtype, np.float64)

    dtimes = np.empty(n)                         # :126-129
    dtimes[:n - 1] = times[1:] - times[:-1]
    dtimes[n - 1] = 1.0

    sigma_obs = np.exp(log_sigma_obs)            # :136
    par_mat = linear_predictor(dat, coeff_fe, coeff_re)   # :143-149
    mu = par_mat[:, 0:n_dim]                     # :152
    tau = np.exp(par_mat[:, n_dim])              # :153
    nu = np.exp(par_mat[:, n_dim + 1])           # :154
    beta = 1 / tau                               # :155
    sigma = 2 * nu / np.sqrt(math.pi * tau)      # :156

    Z = np.zeros((n_dim, 2 * n_dim))             # :162-166
    for i in range(n_dim):
        Z[i, 2 * i] = 1
    H = makeH_ctcrw(sigma_obs, n_dim, dtype)     # :167

    aest = a0[0].astype(dtype)                   # :181-185
    Pest = P0.astype(dtype)
    k = 1
    llk = 0.0
    aest_all = np.zeros((n, 2 * n_dim), dtype=dtype)
    aest_all[0] = aest
    for i in range(1, n):                        # :195
        if ID[i] != ID[i - 1]:                   # :196-200
            aest = a0[k].astype(dtype)
            k += 1
            Pest = P0.astype(dtype)
        else:
            if H_array is not None and np.size(H_array) > 1:   # :203-205
                H = np.asarray(H_array)[:, :, i].astype(dtype)
            T = makeT_ctcrw(beta[i], dtimes[i], n_dim, dtype)             # :206
            Q = makeQ_ctcrw(beta[i], sigma[i], dtimes[i], n_dim, dtype)   # :207
            B = makeB_ctcrw(beta[i], dtimes[i], n_dim, dtype)             # :208
            B_times_mu = B @ mu[i]                                        # :211-212
            if _isna(obs[i, 0]):                 # :214-217  (only column 0 is tested)
                aest = T @ aest + B_times_mu
                Pest = T @ Pest @ T.T + Q
            else:
                u = obs[i] - Z @ aest            # :221
                F = Z @ Pest @ Z.T + H           # :223
                detF = det_small(F)              # :224
                if _re(detF) <= 0:               # :226-228 (note: no B*mu here)
                    aest = T @ aest
                    Pest = T @ Pest @ T.T + Q
                else:
                    Finv = np.linalg.inv(F)
                    FinvTu = Finv.T @ u          # :231-232
                    uFu = np.sum(u * FinvTu)     # :233
                    llk = llk - (np.log(detF) + uFu) / 2      # :234 (no d*log(2*pi))
                    K = T @ Pest @ Z.T @ Finv    # :236
                    aest = T @ aest + K @ u + B_times_mu      # :238
                    L = T - K @ Z                # :240
                    Pest = T @ Pest @ L.T + Q    # :241
        aest_all[i] = aest                       # :246

    nllk = -llk + penalty_kalman(dat, log_lambda, coeff_re)   # :254-280
    if return_aest:
        return nllk, aest_all
    return nllk


# --------------------------------------------------------------------------------------------
# nllk_ou_ssm / nllk_bm_ssm: the same Kalman loop with an n_dim-state filter
# (nllk_ou_ssm.hpp:73-249, nllk_bm_ssm.hpp:40-211)
# --------------------------------------------------------------------------------------------
def makeT_ou_ssm(tau, dt, n_dim, dtype):
    """nllk_ou_ssm.hpp:30-38"""
    return np.eye(n_dim, dtype=dtype) * np.exp(-dt / tau)


def makeB_ou_ssm(tau, dt, n_dim, dtype):
    """nllk_ou_ssm.hpp:45-53"""
    return np.eye(n_dim, dtype=dtype) * (1 - np.exp(-dt / tau))


def makeQ_ou_ssm(tau, kappa, dt, n_dim, dtype):
    """nllk_ou_ssm.hpp:61-69"""
    return np.eye(n_dim, dtype=dtype) * (kappa * (1 - np.exp(-2 * dt / tau)))


def makeQ_bm_ssm(sigma, dt, n_dim, dtype):
    """nllk_bm_ssm.hpp:27-36"""
    return np.eye(n_dim, dtype=dtype) * (sigma * sigma * dt)


def _nllk_ssm(dat, log_sigma_obs, coeff_fe, log_lambda, coeff_re, model, return_aest=False):
    """OU_SSM: nllk_ou_ssm.hpp:101-249; BM_SSM: nllk_bm_ssm.hpp:62-211 (line numbers of OU_SSM below)."""
    ID = np.asarray(dat["ID"])
    times = np.asarray(dat["times"], dtype=float)
    obs = np.asarray(dat["obs"], dtype=float)
    a0 = np.asarray(dat["a0"], dtype=float)
    P0 = np.asarray(dat["P0"], dtype=float)
    H_array = dat.get("H_array", None)
    n, n_dim = obs.shape
    coeff_fe, coeff_re, log_lambda = np.asarray(coeff_fe), np.asarray(coeff_re), np.asarray(log_lambda)
    dtype = np.result_type(coeff_fe.dtype, coeff_re.dtype, np.asarray(log_sigma_obs).dtype, log_lambda.dtype, np.float64)
    dtimes = np.empty(n)                          # :97-100
    dtimes[:n - 1] = times[1:] - times[:-1]
    dtimes[n - 1] = 1.0
    sigma_obs = np.exp(log_sigma_obs)             # :106-107
    par_mat = linear_predictor(dat, coeff_fe, coeff_re)    # :113-119
    mu = par_mat[:, 0:n_dim]                      # :122
    if model == "OU_SSM":
        tau = np.exp(par_mat[:, n_dim])           # :123
        kappa = np.exp(par_mat[:, n_dim + 1])     # :124
    else:
        sigma = np.exp(par_mat[:, n_dim])         # nllk_bm_ssm.hpp:90
    Z = np.eye(n_dim)                             # :130-131
    H = np.eye(n_dim, dtype=dtype) * (sigma_obs * sigma_obs)   # :132, makeH :15-23
    aest = a0[0].astype(dtype)                    # :150-154
    Pest = P0.astype(dtype)
    k = 1
    llk = 0.0
    aest_all = np.zeros((n, n_dim), dtype=dtype)
    aest_all[0] = aest
    for i in range(1, n):                         # :163
        if ID[i] != ID[i - 1]:                    # :164-168
            aest = a0[k].astype(dtype)
            k += 1
            Pest = P0.astype(dtype)
        else:
            if H_array is not None and np.size(H_array) > 1:      # :171-173
                H = np.asarray(H_array)[:, :, i].astype(dtype)
            if model == "OU_SSM":
                T = makeT_ou_ssm(tau[i], dtimes[i], n_dim, dtype)             # :174
                B = makeB_ou_ssm(tau[i], dtimes[i], n_dim, dtype)             # :175
                Q = makeQ_ou_ssm(tau[i], kappa[i], dtimes[i], n_dim, dtype)   # :176
                drift = B @ mu[i]                                             # :177
            else:
                T = np.eye(n_dim, dtype=dtype)                                # nllk_bm_ssm.hpp:100-101
                Q = makeQ_bm_ssm(sigma[i], dtimes[i], n_dim, dtype)           # nllk_bm_ssm.hpp:139
                drift = mu[i] * dtimes[i]                                     # nllk_bm_ssm.hpp:140
            if _isna(obs[i, 0]):                  # :179-182
                aest = T @ aest + drift
                Pest = T @ Pest @ T.T + Q
            else:
                u = obs[i] - Z @ aest             # :185-186
                F = Z @ Pest @ Z.T + H            # :189
                detF = det_small(F)               # detF = exp(atomic::logdet(F)) :190 (analytic form, so that
                                                  # complex-step derivatives pass through it)
                if _re(detF) <= 0:                # :192-194 (keeps B*mu, unlike CTCRW)
                    aest = T @ aest + drift
                    Pest = T @ Pest @ T.T + Q
                else:
                    Finv = np.linalg.inv(F)
                    uFu = np.sum(u * (Finv.T @ u))            # :197-199
                    llk = llk - (np.log(detF) + uFu) / 2      # :200
                    K = T @ Pest @ Z.T @ Finv                 # :202
                    aest = T @ aest + K @ u + drift           # :204
                    L = T - K @ Z                             # :206
                    Pest = T @ Pest @ L.T + Q                 # :207
        aest_all[i] = aest                        # :212
    nllk = -llk + penalty_kalman(dat, log_lambda, coeff_re)   # :220-246 (identical to nllk_ctcrw's)
    if return_aest:
        return nllk, aest_all
    return nllk


def nllk_ou_ssm(dat, log_sigma_obs, coeff_fe, log_lambda, coeff_re, return_aest=False):
    return _nllk_ssm(dat, log_sigma_obs, coeff_fe, log_lambda, coeff_re, "OU_SSM", return_aest)


def nllk_bm_ssm(dat, log_sigma_obs, coeff_fe, log_lambda, coeff_re, return_aest=False):
    return _nllk_ssm(dat, log_sigma_obs, coeff_fe, log_lambda, coeff_re, "BM_SSM", return_aest)


# --------------------------------------------------------------------------------------------
# tr_dens / nllk_sde, tr_dens.hpp:18-75 and nllk_sde.hpp:15-127  (BM and OU only)
# --------------------------------------------------------------------------------------------
def dnorm_log(x, mean, sd):
    """TMB dnorm(x, mean, sd, true): -log(sqrt(2*pi)) - log(sd) - 0.5*((x-mean)/sd)^2."""
    resid = (x - mean) / sd
    return -LOG_SQRT_2PI - np.log(sd) - 0.5 * resid * resid


def tr_dens(Z1, Z0, dtimes, par, type_):
    """tr_dens.hpp:18-75 with give_log = true; BM (:32-37) and OU (:45-52) branches."""
    n_dim = Z1.shape[0]
    res = 0.0
    for i in range(n_dim):
        if (not _isna(Z0[i])) and (not _isna(Z1[i])):      # :31
            if type_ == "BM":
                mean = Z0[i] + par[i] * dtimes              # :35
                sd = np.exp(par[n_dim]) * np.sqrt(dtimes)   # :36
                res = res + dnorm_log(Z1[i], mean, sd)
            elif type_ == "OU":
                mean = par[i] + np.exp(-dtimes / np.exp(par[n_dim])) * (Z0[i] - par[i])       # :49
                sd = np.sqrt(np.exp(par[n_dim + 1]) * (1 - np.exp(-2 * dtimes / np.exp(par[n_dim]))))
                res = res + dnorm_log(Z1[i], mean, sd)      # :50-52
            else:
                raise ValueError("oracle covers BM and OU only (BM_t / CIR are out of scope)")
    return res


def has_decay(dat):
    """t_decay.size() > 1, nllk_sde.hpp:49 (R passes t_decay = 0 when there is no decay term, R/sde.R:645)."""
    t = dat.get("t_decay")
    return t is not None and np.size(t) > 1


def n_decay(dat):
    """length(log_decay) = length(unique(ind_decay)), R/sde.R:177,506."""
    return int(np.unique(np.asarray(dat["ind_decay"])).size) if has_decay(dat) else 0


def decayed_X_re(dat, log_decay):
    """nllk_sde.hpp:47-59: X_re_copy = X_re with column col_decay(i) multiplied, row by row, by
    exp(-exp(log_decay(ind_decay(i))) * t_decay); col_decay / ind_decay are 1-based."""
    import scipy.sparse as sp
    X = sp.csc_matrix(dat["X_re"]).astype(np.asarray(log_decay).dtype)
    decay_rate = np.exp(np.asarray(log_decay))                    # :47
    t_decay = np.asarray(dat["t_decay"], dtype=float)
    X = sp.lil_matrix(X)
    Xc = sp.csc_matrix(dat["X_re"])
    for i_col, i_ind in zip(np.atleast_1d(dat["col_decay"]), np.atleast_1d(dat["ind_decay"])):
        c, k = int(i_col) - 1, int(i_ind) - 1                     # :51-52
        decay = np.exp(-decay_rate[k] * t_decay)                  # :53
        col = Xc[:, c]
        rows = col.indices
        X[rows, c] = (np.asarray(col.data) * decay[rows]).reshape(-1, 1)      # :54-56
    return sp.csr_matrix(X)


def nllk_sde(dat, coeff_fe, log_lambda, coeff_re, type_=None, log_decay=None):
    type_ = type_ or dat["type"]
    if has_decay(dat):
        dat = dict(dat, X_re=decayed_X_re(dat, log_decay))
    ID = np.asarray(dat["ID"])
    times = np.asarray(dat["times"], dtype=float)
    obs = np.asarray(dat["obs"], dtype=float)
    n = obs.shape[0]
    dtimes = np.diff(times)                      # nllk_sde.hpp:37
    coeff_fe = np.asarray(coeff_fe)
    coeff_re = np.asarray(coeff_re)
    log_lambda = np.asarray(log_lambda)
    par_mat = linear_predictor(dat, coeff_fe, coeff_re)   # :61-67
    llk = 0.0
    for i in range(1, n):                        # :77-84
        if ID[i - 1] == ID[i]:
            llk = llk + tr_dens(obs[i], obs[i - 1], dtimes[i - 1], par_mat[i - 1], type_)
    return -llk + penalty_sde(dat, log_lambda, coeff_re)  # :89-124


# --------------------------------------------------------------------------------------------
# dispatch, smoothSDE.cpp:9-28, on the flat parameter vector of SURVEY.md section 8(a) row A1
# --------------------------------------------------------------------------------------------
def split_par(dat, par):
    """Flat joint parameter vector -> named pieces.
    CTCRW: [log_sigma_obs, coeff_fe, log_lambda, coeff_re]  (nllk_ctcrw.hpp:135-140)
    BM/OU: [coeff_fe, log_lambda, (log_decay,) coeff_re]  (nllk_sde.hpp:42-45; log_decay is mapped
    off whenever no decay term exists, R/sde.R:648, and then absent from the flat vector)."""
    par = np.asarray(par)
    p_fe = dat["X_fe"].shape[1]
    p_re = dat["X_re"].shape[1]
    ncol_re = np.atleast_1d(np.asarray(dat["ncol_re"], dtype=np.int64))
    n_s = ncol_re.size if ncol_re[0] > 0 else 1
    o = 0
    out = {}
    if dat["type"] in KALMAN_TYPES:
        out["log_sigma_obs"] = par[0]
        o = 1
    out["coeff_fe"] = par[o:o + p_fe]; o += p_fe
    out["log_lambda"] = par[o:o + n_s]; o += n_s
    if dat["type"] not in KALMAN_TYPES and has_decay(dat):
        nd_ = n_decay(dat)
        out["log_decay"] = par[o:o + nd_]; o += nd_
    out["coeff_re"] = par[o:o + p_re]; o += p_re
    assert o == par.size, (o, par.size)
    return out


def nllk(dat, par):
    p = split_par(dat, par)
    t = dat["type"]
    if t in ("BM", "OU"):
        return nllk_sde(dat, p["coeff_fe"], p["log_lambda"], p["coeff_re"], t, log_decay=p.get("log_decay"))
    if t == "CTCRW":
        return nllk_ctcrw(dat, p["log_sigma_obs"], p["coeff_fe"], p["log_lambda"], p["coeff_re"])
    if t in ("OU_SSM", "BM_SSM"):
        return _nllk_ssm(dat, p["log_sigma_obs"], p["coeff_fe"], p["log_lambda"], p["coeff_re"], t)
    raise ValueError("Unknown SDE type")         # smoothSDE.cpp:25


def grad_complex_step(dat, par, h=1e-30):
    """Complex-step gradient of nllk(dat, .) -- exact to rounding for analytic objectives."""
    par = np.asarray(par, dtype=float)
    g = np.empty(par.size)
    for j in range(par.size):
        z = par.astype(complex)
        z[j] += 1j * h
        g[j] = np.imag(nllk(dat, z)) / h
    return g


def hess_complex_fd(dat, par, k=1e-3, h=1e-30):
    """Hessian of nllk(dat, .): complex-step first derivative (exact to rounding), Richardson-
    extrapolated central difference of it in the second direction (error O(k^4))."""
    par = np.asarray(par, dtype=float)
    n = par.size
    H = np.empty((n, n))

    def g(p):
        return grad_complex_step(dat, p, h)

    for j in range(n):
        e = np.zeros(n)
        e[j] = 1.0
        d1 = (g(par + k * e) - g(par - k * e)) / (2 * k)
        d2 = (g(par + 0.5 * k * e) - g(par - 0.5 * k * e)) / k
        H[:, j] = (4 * d2 - d1) / 3
    return 0.5 * (H + H.T)
