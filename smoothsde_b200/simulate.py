"""Exact-transition simulators for BM / OU / CTCRW tracks (host side, numpy).

Specification followed: SDE$simulate, R/sde.R:1395-1508 (BM :1434-1438, OU :1439-1447, CTCRW
:1448-1478) and CTCRW_cov, R/utility.R:188-196.  Parameters are natural-scale arrays with one
row per observation (the R code takes them from ``self$par(new_data)``); the transition from row
i-1 to row i uses the parameters of row i-1, as in the reference.

Tracks are simulated side by side (vectorised over tracks, sequential over time) which is what
makes the 10^5-step-per-track configurations affordable in numpy.
"""
from __future__ import annotations

import numpy as np


def ctcrw_cov(beta, sigma, dt):
    """R/utility.R:188-196, (velocity, position) ordering; returns (Qvv, Qzz, Qvz)."""
    e = np.exp(-beta * dt)
    e2 = np.exp(-2 * beta * dt)
    qvv = sigma ** 2 / (2 * beta) * (1 - e2)
    qzz = (sigma / beta) ** 2 * (dt + (1 - e2) / (2 * beta) - 2 * (1 - e) / beta)
    qvz = sigma ** 2 / (2 * beta ** 2) * (1 - 2 * e + e2)
    return qvv, qzz, qvz


def make_times(n_tracks, n_steps, rng, irregular=False):
    """[n_tracks, n_steps] observation times: t = 0,1,2,... or cumsum(U(0.2, 2.0))."""
    if irregular:
        inc = rng.uniform(0.2, 2.0, size=(n_tracks, n_steps))
        inc[:, 0] = 0.0
        return np.cumsum(inc, axis=1)
    return np.tile(np.arange(n_steps, dtype=float), (n_tracks, 1))


def simulate_bm(times, mu, sigma, rng, z0=0.0):
    """times [T, m]; mu, sigma [T, m] natural scale.  R/sde.R:1434-1438."""
    dt = np.diff(times, axis=1)
    inc = rng.normal(mu[:, :-1] * dt, sigma[:, :-1] * np.sqrt(dt))
    z = np.concatenate([np.full((times.shape[0], 1), z0), z0 + np.cumsum(inc, axis=1)], axis=1)
    return z


def simulate_ou(times, mu, tau, kappa, rng, z0=0.0):
    """R/sde.R:1439-1447."""
    T, m = times.shape
    z = np.empty((T, m))
    z[:, 0] = z0
    dt = np.diff(times, axis=1)
    for i in range(1, m):
        p = np.exp(-dt[:, i - 1] / tau[:, i - 1])
        mean = p * z[:, i - 1] + (1 - p) * mu[:, i - 1]
        sd = np.sqrt(kappa[:, i - 1] * (1 - p * p))
        z[:, i] = rng.normal(mean, sd)
    return z


def simulate_ctcrw(times, mu, tau, nu, rng, z0=0.0):
    """One spatial dimension of a CTCRW; call once per dimension.  R/sde.R:1448-1478."""
    T, m = times.shape
    beta = 1.0 / tau
    sigma = 2.0 * nu / np.sqrt(np.pi * tau)
    z = np.empty((T, m))
    v = np.zeros(T)
    z[:, 0] = z0
    dt = np.diff(times, axis=1)
    for i in range(1, m):
        b, s, d, mm = beta[:, i - 1], sigma[:, i - 1], dt[:, i - 1], mu[:, i - 1]
        p = np.exp(-b * d)
        mean_v = p * v + (1 - p) * mm
        mean_z = z[:, i - 1] + mm * d + (v - mm) / b * (1 - p)
        qvv, qzz, qvz = ctcrw_cov(b, s, d)
        # Cholesky of [[qvv, qvz], [qvz, qzz]]
        l11 = np.sqrt(qvv)
        l21 = qvz / l11
        l22 = np.sqrt(np.maximum(qzz - l21 * l21, 0.0))
        e1 = rng.standard_normal(T)
        e2 = rng.standard_normal(T)
        v = mean_v + l11 * e1
        z[:, i] = mean_z + l21 * e1 + l22 * e2
    return z
