"""Independent known-answer checks that pin the oracle (TEST INFRASTRUCTURE ONLY).

The reference's test-suite holds no numeric vector for the likelihood path and TMB cannot be run
in this image, so the numpy restatement oracle_np.py is pinned by computing the
same numbers in completely different ways:

  * BM / OU (nllk_sde.hpp:77-84 + tr_dens.hpp:32-37,45-52): the objective is minus a sum of
    univariate normal log-densities -> scipy.stats.norm.logpdf, vectorised over transitions;
  * CTCRW (nllk_ctcrw.hpp:195-247): the Kalman filter's prediction-error decomposition equals
    the log-density of the stacked observations of a track under the linear-Gaussian state-space
    model, with the prior N(a0, P0) sitting on the state of the track's SECOND row
     -> one dense multivariate normal per track, no recursion at all;
    the template drops the d*log(2 pi)/2 constant of every observed row (:231-234), which is
    added back here;
  * nllk_mpmath: the recursions re-evaluated with 40 significant digits, which bounds the
    float64 oracle's own rounding error.

Nothing in the product path imports this module.
"""
from __future__ import annotations

import math

import numpy as np
import scipy.sparse as sp
from scipy.stats import multivariate_normal, norm

from . import oracle_np as O


def _par_mat(dat, p):
    return O.linear_predictor(dat, p["coeff_fe"], p["coeff_re"])


def known_sde(dat, par):
    p = O.split_par(dat, par)
    pm = _par_mat(dat, p)
    ID, t, obs = np.asarray(dat["ID"]), np.asarray(dat["times"], float), np.asarray(dat["obs"], float)
    n, d = obs.shape
    same = ID[1:] == ID[:-1]
    dt = np.diff(t)
    total = 0.0
    for k in range(d):
        z0, z1 = obs[:-1, k], obs[1:, k]
        ok = same & ~np.isnan(z0) & ~np.isnan(z1)
        if dat["type"] == "BM":
            mean = z0 + pm[:-1, k] * dt
            sd = np.exp(pm[:-1, d]) * np.sqrt(dt)
        else:
            tau, kappa = np.exp(pm[:-1, d]), np.exp(pm[:-1, d + 1])
            ph = np.exp(-dt / tau)
            mean = pm[:-1, k] + ph * (z0 - pm[:-1, k])
            sd = np.sqrt(kappa * (1 - np.exp(-2 * dt / tau)))
        total += norm.logpdf(z1[ok], mean[ok], sd[ok]).sum()
    return -total + O.penalty_sde(dat, p["log_lambda"], p["coeff_re"])


def known_ctcrw(dat, par):
    p = O.split_par(dat, par)
    pm = _par_mat(dat, p)
    ID, t, obs = np.asarray(dat["ID"]), np.asarray(dat["times"], float), np.asarray(dat["obs"], float)
    a0, P0 = np.asarray(dat["a0"], float), np.asarray(dat["P0"], float)
    n, d = obs.shape
    m2 = 2 * d
    h = math.exp(2 * float(p["log_sigma_obs"]))
    tau, nu = np.exp(pm[:, d]), np.exp(pm[:, d + 1])
    beta, sigma = 1 / tau, 2 * nu / np.sqrt(math.pi * tau)
    Z = np.zeros((d, m2))
    Z[np.arange(d), 2 * np.arange(d)] = 1
    starts = np.r_[0, np.nonzero(ID[1:] != ID[:-1])[0] + 1, n]
    llk = 0.0
    for k in range(starts.size - 1):
        s, e = starts[k], starts[k + 1]
        rows = np.arange(s + 1, e)              # rows that can contribute
        q = rows.size
        if q == 0:
            continue
        # joint distribution of the states x_{s+1..e-1}: x_{s+1} ~ N(a0, P0),
        # x_{i+1} = T_i x_i + B_i mu_i + N(0, Q_i) with row i's parameters and dt_i = t_{i+1}-t_i
        mean = np.zeros((q, m2))
        cov = np.zeros((q, q, m2, m2))
        mean[0] = a0[k]
        cov[0, 0] = P0
        for j in range(1, q):
            i = rows[j - 1]
            dti = t[i + 1] - t[i]
            T = O.makeT_ctcrw(beta[i], dti, d, float)
            Q = O.makeQ_ctcrw(beta[i], sigma[i], dti, d, float)
            B = O.makeB_ctcrw(beta[i], dti, d, float)
            mean[j] = T @ mean[j - 1] + B @ pm[i, :d]
            for l in range(j):
                cov[j, l] = T @ cov[j - 1, l]
                cov[l, j] = cov[j, l].T
            cov[j, j] = T @ cov[j - 1, j - 1] @ T.T + Q
        seen = np.nonzero(~np.isnan(obs[rows, 0]))[0]
        if seen.size == 0:
            continue
        my = (mean[seen] @ Z.T).ravel()
        Cy = np.zeros((seen.size * d, seen.size * d))
        for a, ja in enumerate(seen):
            for b_, jb in enumerate(seen):
                Cy[a * d:(a + 1) * d, b_ * d:(b_ + 1) * d] = Z @ cov[ja, jb] @ Z.T
        H_array = dat.get("H_array")
        if H_array is not None and np.size(H_array) > 1:        # user measurement covariances, nllk_ctcrw.hpp:203-205
            for a, ja in enumerate(seen):
                Cy[a * d:(a + 1) * d, a * d:(a + 1) * d] += np.asarray(H_array, float)[:, :, rows[ja]]
        else:
            Cy += h * np.eye(seen.size * d)
        y = obs[rows[seen]].ravel()
        llk += multivariate_normal(my, Cy, allow_singular=False).logpdf(y)
        llk += 0.5 * seen.size * d * math.log(2 * math.pi)      # constant the template omits
    return -llk + O.penalty_kalman(dat, p["log_lambda"], p["coeff_re"])


def known_ssm(dat, par):
    """OU_SSM / BM_SSM: one dense multivariate-normal density per track and dimension.  The state of
    a track's second row has prior N(a0, P0); x_{i+1} = t_i x_i + c_i + N(0, q_i); y_i = x_i + N(0, h)."""
    p = O.split_par(dat, np.asarray(par, float))
    pm = _par_mat(dat, p)
    ID, t, obs = np.asarray(dat["ID"]), np.asarray(dat["times"], float), np.asarray(dat["obs"], float)
    a0, P0 = np.asarray(dat["a0"], float), np.asarray(dat["P0"], float)
    n, d = obs.shape
    h = math.exp(2 * float(p["log_sigma_obs"]))
    starts = np.r_[0, np.nonzero(ID[1:] != ID[:-1])[0] + 1, n]
    llk = 0.0
    for k in range(starts.size - 1):
        s, e = starts[k], starts[k + 1]
        rows = np.arange(s + 1, e)
        q = rows.size
        if q == 0:
            continue
        tt, qq, cc = np.ones(q), np.zeros(q), np.zeros((q, d))
        for j in range(1, q):
            i = rows[j - 1]
            dti = t[i + 1] - t[i]
            if dat["type"] == "OU_SSM":
                tau, kappa = math.exp(pm[i, d]), math.exp(pm[i, d + 1])
                tt[j], qq[j] = math.exp(-dti / tau), kappa * (1 - math.exp(-2 * dti / tau))
                cc[j] = (1 - tt[j]) * pm[i, :d]
            else:
                tt[j], qq[j] = 1.0, math.exp(2 * pm[i, d]) * dti
                cc[j] = pm[i, :d] * dti
        seen = np.nonzero(~np.isnan(obs[rows, 0]))[0]
        if seen.size == 0:
            continue
        for dim in range(d):
            mean = np.zeros(q)
            cov = np.zeros((q, q))
            mean[0], cov[0, 0] = a0[k, dim], P0[dim, dim]
            for j in range(1, q):
                mean[j] = tt[j] * mean[j - 1] + cc[j, dim]
                cov[j, :j] = tt[j] * cov[j - 1, :j]
                cov[:j, j] = cov[j, :j]
                cov[j, j] = tt[j] ** 2 * cov[j - 1, j - 1] + qq[j]
            Cy = cov[np.ix_(seen, seen)] + h * np.eye(seen.size)
            llk += multivariate_normal(mean[seen], Cy, allow_singular=False).logpdf(obs[rows[seen], dim])
        llk += 0.5 * seen.size * d * math.log(2 * math.pi)      # constant the templates omit
    return -llk + O.penalty_kalman(dat, p["log_lambda"], p["coeff_re"])


def known_answer(dat, par):
    if dat["type"] == "CTCRW":
        return known_ctcrw(dat, par)
    if dat["type"] in ("OU_SSM", "BM_SSM"):
        return known_ssm(dat, par)
    return known_sde(dat, par)


# --------------------------------------------------------------------------------------------
# 40-digit evaluation
# --------------------------------------------------------------------------------------------
def nllk_mpmath(dat, par, dps=40):
    import mpmath as mp
    mp.mp.dps = dps
    p = O.split_par(dat, np.asarray(par, float))
    obs = np.asarray(dat["obs"], float)
    n, d = obs.shape
    ID, t = np.asarray(dat["ID"]), np.asarray(dat["times"], float)
    X = sp.hstack([sp.csr_matrix(dat["X_fe"]), sp.csr_matrix(dat["X_re"])], format="csr")
    theta = [mp.mpf(float(x)) for x in np.concatenate([p["coeff_fe"], p["coeff_re"]])]
    pv = []
    for r in range(X.shape[0]):
        acc = mp.mpf(0)
        for kk in range(X.indptr[r], X.indptr[r + 1]):
            acc += mp.mpf(float(X.data[kk])) * theta[X.indices[kk]]
        pv.append(acc)
    n_par = X.shape[0] // n
    pm = [[pv[j * n + i] for j in range(n_par)] for i in range(n)]
    typ = dat["type"]
    llk = mp.mpf(0)
    if typ in ("BM", "OU"):
        for i in range(1, n):
            if ID[i] != ID[i - 1]:
                continue
            dt = mp.mpf(float(t[i])) - mp.mpf(float(t[i - 1]))
            q = pm[i - 1]
            for k in range(d):
                if np.isnan(obs[i, k]) or np.isnan(obs[i - 1, k]):
                    continue
                z0, z1 = mp.mpf(float(obs[i - 1, k])), mp.mpf(float(obs[i, k]))
                if typ == "BM":
                    mean = z0 + q[k] * dt
                    sd = mp.exp(q[d]) * mp.sqrt(dt)
                else:
                    mean = q[k] + mp.exp(-dt / mp.exp(q[d])) * (z0 - q[k])
                    sd = mp.sqrt(mp.exp(q[d + 1]) * (1 - mp.exp(-2 * dt / mp.exp(q[d]))))
                r = (z1 - mean) / sd
                llk += -mp.log(mp.sqrt(2 * mp.pi)) - mp.log(sd) - r * r / 2
        pen = O.penalty_sde(dat, p["log_lambda"], p["coeff_re"])
        return float(-llk + mp.mpf(pen))
    # CTCRW: dense recursion of nllk_ctcrw.hpp:181-247 in mp.matrix arithmetic
    m2 = 2 * d
    a0, P0 = np.asarray(dat["a0"], float), np.asarray(dat["P0"], float)
    h = mp.exp(2 * mp.mpf(float(p["log_sigma_obs"])))
    Z = mp.zeros(d, m2)
    for k in range(d):
        Z[k, 2 * k] = 1
    H = mp.eye(d) * h
    a = mp.matrix([float(x) for x in a0[0]])
    P = mp.matrix(P0.tolist())
    ktrack = 1
    for i in range(1, n):
        if ID[i] != ID[i - 1]:
            a = mp.matrix([float(x) for x in a0[ktrack]])
            ktrack += 1
            P = mp.matrix(P0.tolist())
            continue
        dt = (mp.mpf(float(t[i + 1])) - mp.mpf(float(t[i]))) if i < n - 1 else mp.mpf(1)
        tau, nu = mp.exp(pm[i][d]), mp.exp(pm[i][d + 1])
        beta, sigma = 1 / tau, 2 * nu / mp.sqrt(mp.pi * tau)
        e1, e2 = mp.exp(-beta * dt), mp.exp(-2 * beta * dt)
        T, Q, B = mp.zeros(m2, m2), mp.zeros(m2, m2), mp.zeros(m2, d)
        for k in range(d):
            T[2 * k, 2 * k] = 1
            T[2 * k, 2 * k + 1] = (1 - e1) / beta
            T[2 * k + 1, 2 * k + 1] = e1
            Q[2 * k, 2 * k] = (sigma / beta) ** 2 * (dt - 2 / beta * (1 - e1) + 1 / (2 * beta) * (1 - e2))
            Q[2 * k, 2 * k + 1] = Q[2 * k + 1, 2 * k] = sigma ** 2 / (2 * beta ** 2) * (1 - 2 * e1 + e2)
            Q[2 * k + 1, 2 * k + 1] = sigma ** 2 / (2 * beta) * (1 - e2)
            B[2 * k, k] = dt - (1 - e1) / beta
            B[2 * k + 1, k] = 1 - e1
        Bmu = B * mp.matrix([pm[i][k] for k in range(d)])
        if np.isnan(obs[i, 0]):
            a = T * a + Bmu
            P = T * P * T.T + Q
            continue
        u = mp.matrix([float(x) for x in obs[i]]) - Z * a
        F = Z * P * Z.T + H
        Fi = F ** -1
        llk -= (mp.log(mp.det(F)) + (u.T * Fi * u)[0]) / 2
        K = T * P * Z.T * Fi
        a = T * a + K * u + Bmu
        P = T * P * (T - K * Z).T + Q
    pen = O.penalty_kalman(dat, p["log_lambda"], p["coeff_re"])
    return float(-llk + mp.mpf(float(pen)))
