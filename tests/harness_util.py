"""Helpers shared by the CPU-side tests: row flags / dt conventions and the ctypes binding of the
host-compiled math harness (tests/harness/math_harness.cpp).  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
_c_dp = ctypes.POINTER(ctypes.c_double)


def _P(a):
    return a.ctypes.data_as(_c_dp)


def build_harness():
    src = os.path.join(HERE, "harness", "math_harness.cpp")
    out_dir = os.path.join(HERE, "harness", "_build")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, "libmath_harness.so")
    hdrs = [os.path.join(ROOT, "smoothsde_b200", "csrc", f) for f in ("ctcrw_math.cuh", "dual.cuh", "ssm1_math.cuh", "models.cuh", "dense_math.cuh")]
    if (not os.path.exists(out)) or os.path.getmtime(out) < max([os.path.getmtime(src)] + [os.path.getmtime(f) for f in hdrs]):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", "-o", out, src])
    return ctypes.CDLL(out)


def row_flags(ID, obs):
    ID = np.asarray(ID)
    start = np.r_[True, ID[1:] != ID[:-1]]
    last = np.r_[ID[1:] != ID[:-1], True]
    has = ~np.isnan(obs[:, 0])
    return (start * 1 + last * 2 + has * 4).astype(np.uint8)


def ctcrw_dt(times, flags):
    """dt_i = t_{i+1} - t_i (nllk_ctcrw.hpp:126-129); rows whose prediction is discarded (last
    row of a track) get dt = 1 like the reference's final row."""
    t = np.asarray(times, dtype=float)
    dt = np.r_[t[1:] - t[:-1], 1.0]
    dt[(flags & 2) != 0] = 1.0
    return dt


def harness_ctcrw(lib, dat, eta, log_sigma_obs, mode, lc=8, nt=128, want_grad=True, want_aest=False):
    obs = np.asarray(dat["obs"], dtype=float)
    n, nd = obs.shape
    flags = row_flags(dat["ID"], obs)
    dt = ctcrw_dt(dat["times"], flags)
    eta = np.ascontiguousarray(eta, dtype=float)
    y = np.ascontiguousarray(np.nan_to_num(obs))
    a0 = np.ascontiguousarray(dat["a0"], dtype=float)
    P0 = np.asarray(dat["P0"], dtype=float)
    P0s = np.array([P0[0, 0], P0[0, 1], P0[1, 1]])
    h = float(np.exp(2 * log_sigma_obs))
    llk = ctypes.c_double()
    gh = ctypes.c_double()
    eb = np.zeros((n, nd + 2)) if want_grad else None
    aest = np.zeros((n, 2 * nd)) if want_aest else None
    rc = lib.harness_ctcrw(nd, mode, ctypes.c_int64(n),
                           flags.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), _P(y), _P(dt),
                           _P(eta), _P(a0), _P(P0s), ctypes.c_double(h), lc, nt,
                           ctypes.byref(llk), _P(eb) if want_grad else None, ctypes.byref(gh),
                           _P(aest) if want_aest else None)
    assert rc == 0
    return llk.value, eb, gh.value * 2 * h, aest


def harness_ctcrw_tangent(lib, dat, eta, eta_dot, log_sigma_obs, lso_dot, mode, lc=8, nt=128):
    """Dual-number run of the scan algebra along (eta_dot, d log_sigma_obs = lso_dot).
    Returns (llk, d llk), (eta_bar, d eta_bar), (d nllk/d log_sigma_obs, its tangent)."""
    obs = np.asarray(dat["obs"], dtype=float)
    n, nd = obs.shape
    flags = row_flags(dat["ID"], obs)
    dt = ctcrw_dt(dat["times"], flags)
    eta = np.ascontiguousarray(eta, dtype=float)
    eta_dot = np.ascontiguousarray(eta_dot, dtype=float)
    y = np.ascontiguousarray(np.nan_to_num(obs))
    a0 = np.ascontiguousarray(dat["a0"], dtype=float)
    P0 = np.asarray(dat["P0"], dtype=float)
    P0s = np.array([P0[0, 0], P0[0, 1], P0[1, 1]])
    h = float(np.exp(2 * log_sigma_obs))
    h_dot = 2 * h * lso_dot
    llk2 = np.zeros(2)
    gh2 = np.zeros(2)
    eb = np.zeros((n, nd + 2))
    ebd = np.zeros((n, nd + 2))
    rc = lib.harness_ctcrw_tangent(nd, mode, ctypes.c_int64(n),
                                   flags.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), _P(y), _P(dt),
                                   _P(eta), _P(eta_dot), _P(a0), _P(P0s), ctypes.c_double(h),
                                   ctypes.c_double(h_dot), lc, nt, _P(llk2), _P(eb), _P(ebd), _P(gh2))
    assert rc == 0
    # g_lso = 2 h gh  ->  d g_lso = 2 (h_dot gh + h gh_dot)
    g_lso = 2 * h * gh2[0]
    g_lso_dot = 2 * (h_dot * gh2[0] + h * gh2[1])
    return llk2, (eb, ebd), (g_lso, g_lso_dot)


MODEL_IDS = {"CTCRW": 0, "OU_SSM": 1, "BM_SSM": 2}


def harness_kalman(lib, dat, eta, log_sigma_obs, mode, lc=8, nt=128):
    """Any Kalman model through the host-compiled traits (models.cuh): (llk, eta_bar, d nllk / d log_sigma_obs)."""
    obs = np.asarray(dat["obs"], dtype=float)
    n, nd = obs.shape
    flags = row_flags(dat["ID"], obs)
    dt = ctcrw_dt(dat["times"], flags)
    eta = np.ascontiguousarray(eta, dtype=float)
    y = np.ascontiguousarray(np.nan_to_num(obs))
    a0 = np.ascontiguousarray(dat["a0"], dtype=float)
    P0 = np.asarray(dat["P0"], dtype=float)
    P0s = np.array([P0[0, 0], P0[0, 1] if P0.shape[0] > 1 else 0.0, P0[1, 1] if P0.shape[0] > 1 else P0[0, 0]])
    if dat["type"] != "CTCRW":
        P0s = np.array([P0[0, 0], 0.0, P0[0, 0]])
    h = float(np.exp(2 * log_sigma_obs))
    llk = ctypes.c_double()
    gh = ctypes.c_double()
    eb = np.zeros(eta.shape)
    rc = lib.harness_kalman(MODEL_IDS[dat["type"]], nd, mode, ctypes.c_int64(n),
                            flags.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), _P(y), _P(dt),
                            _P(eta), _P(a0), _P(P0s), ctypes.c_double(h), lc, nt,
                            ctypes.byref(llk), _P(eb), ctypes.byref(gh), None)
    assert rc == 0
    return llk.value, eb, gh.value * 2 * h


def pack_H_planes(H_array):
    """H_array [d, d, n] -> [d(d+1)/2, n] planes of the packed upper triangles (row-major order)."""
    Ha = np.asarray(H_array, dtype=float)
    d = Ha.shape[0]
    return np.ascontiguousarray(np.stack([0.5 * (Ha[r, c] + Ha[c, r]) for r in range(d) for c in range(r, d)]))


def harness_dense(lib, dat, eta, log_sigma_obs, mode, lc=8, nt=64, eta_dot=None, lso_dot=0.0, want_aest=False):
    """Coupled filter (DenseModel in models.cuh) through the host-compiled algebra.  Uses dat["P0"]
    as a full matrix and dat.get("H_array").  Returns (llk, eta_bar, d nllk / d log_sigma_obs, aest)
    or, with eta_dot, ((llk, dllk), (eta_bar, eta_bar_dot), (g_lso, g_lso_dot))."""
    obs = np.asarray(dat["obs"], dtype=float)
    n, nd = obs.shape
    flags = row_flags(dat["ID"], obs)
    dt = ctcrw_dt(dat["times"], flags)
    eta = np.ascontiguousarray(eta, dtype=float)
    y = np.ascontiguousarray(np.nan_to_num(obs))
    a0 = np.ascontiguousarray(dat["a0"], dtype=float)
    P0 = np.ascontiguousarray(dat["P0"], dtype=float)
    m = P0.shape[0]
    Hp = None
    if dat.get("H_array") is not None and np.size(dat["H_array"]) > 1:
        Hp = pack_H_planes(dat["H_array"])
    h = float(np.exp(2 * log_sigma_obs))
    h_dot = 2 * h * lso_dot
    llk2, gh2 = np.zeros(2), np.zeros(2)
    eb = np.zeros(eta.shape)
    ebd = np.zeros(eta.shape) if eta_dot is not None else None
    aest = np.zeros((n, m)) if want_aest else None
    ed = np.ascontiguousarray(eta_dot, dtype=float) if eta_dot is not None else None
    rc = lib.harness_dense(MODEL_IDS[dat["type"]], nd, mode, ctypes.c_int64(n),
                           flags.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), _P(y), _P(dt), _P(eta),
                           _P(ed) if ed is not None else None, _P(a0), _P(P0), m, _P(Hp) if Hp is not None else None,
                           ctypes.c_double(h), ctypes.c_double(h_dot), lc, nt, _P(llk2), _P(eb),
                           _P(ebd) if ebd is not None else None, _P(gh2), _P(aest) if want_aest else None)
    assert rc == 0
    if eta_dot is None:
        return llk2[0], eb, gh2[0] * 2 * h, aest
    return llk2, (eb, ebd), (2 * h * gh2[0], 2 * (h_dot * gh2[0] + h * gh2[1]))
