"""Device-side synthetic problems at BASELINE.json's full sizes (10^8 rows do not fit a host
triplet list), built with torch directly in the engine's packed layout and handed to
``ssde_create_packed`` without a copy.  torch is plumbing here (device memory + RNG); the
likelihood itself is computed by the CUDA library.

Shapes follow SURVEY.md section 8(d): CTCRW, d = 2, ``tau, nu ~ s(time, k = 10)``, mu fixed at 0
(C3: 1024 x 1e5 irregular steps; C4: one 1e8-step track; C5: 4096 x 2.5e4).  Every row has the
same 22 nonzeros: mu1.(Intercept), mu2.(Intercept), tau.(Intercept) + 9 spline columns,
nu.(Intercept) + 9 spline columns.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import scipy.sparse as sp

from . import _lib as L
from . import design as _design
from .engine import Engine, _as_triplet


def _bspline_torch(x, k, lo, hi):
    import torch
    nseg = k - 3
    h = (hi - lo) / nseg
    t = (x - lo) / h
    seg = torch.clamp(torch.floor(t), 0, nseg - 1)
    u = t - seg
    seg = seg.to(torch.int64)
    w = torch.stack([(1 - u) ** 3 / 6.0,
                     (3 * u ** 3 - 6 * u ** 2 + 4) / 6.0,
                     (-3 * u ** 3 + 3 * u ** 2 + 3 * u + 1) / 6.0,
                     u ** 3 / 6.0], dim=1)
    B = torch.zeros((x.numel(), k), dtype=torch.float64, device=x.device)
    idx = seg[:, None] + torch.arange(4, device=x.device)[None, :]
    B.scatter_(1, idx, w)
    return B


def simulate_ctcrw_device(times, tau, nu, e1, e2, device):
    """Exact-transition CTCRW simulation (R/sde.R:1448-1478), one thread per track, by the
    library's own simulator kernel (ssde_simulate_ctcrw).  times/tau/nu and the standard normal
    draws e1/e2: [T, m] contiguous.  Returns z [T, m] for one dimension (mu = 0, start at 0)."""
    import torch
    T, m = times.shape
    z = torch.zeros((T, m), dtype=torch.float64, device=times.device)
    lib = L.load()
    torch.cuda.synchronize(times.device)
    rc = lib.ssde_simulate_ctcrw(device, T, m, times.data_ptr(), tau.data_ptr(), nu.data_ptr(), None,
                                 e1.data_ptr(), e2.data_ptr(), z.data_ptr(), None)
    if rc != 0:
        raise L.EngineError(rc, lib.ssde_create_error().decode())
    torch.cuda.synchronize(times.device)
    return z


def permute_rows(x, n_pad, lc, fill):
    """[n, ...] natural row order -> [n_pad, ...] in the engine's warp-tile order
    (row q*WT + l*LC + k  ->  position q*WT + k*32 + l)."""
    import torch
    n = x.shape[0]
    if n_pad > n:
        pad = torch.full((n_pad - n,) + tuple(x.shape[1:]), fill, dtype=x.dtype, device=x.device)
        x = torch.cat([x, pad], dim=0)
    rest = tuple(x.shape[1:])
    x = x.reshape((n_pad // (32 * lc), 32, lc) + rest).transpose(1, 2)
    return x.contiguous().reshape((n_pad,) + rest)


RNG_GROUPS = 16


class GroupedRNG:
    """Random draws that do not depend on how the tracks are spread over ranks: the `total` tracks
    (or track segments) are cut into RNG_GROUPS groups of consecutive tracks, every group has its
    own generator seeded by (seed, global group index), and a rank draws for the groups it holds.
    So bench.py evaluates the SAME data set at 1, 2, 4 and 8 GPUs."""

    def __init__(self, seed, total, local, rank, dev):
        import torch
        self.g = total // RNG_GROUPS if total % RNG_GROUPS == 0 else 1
        if local % self.g != 0:
            raise ValueError(f"{local} tracks per rank is not a multiple of the RNG group size {self.g}")
        first = rank * local // self.g
        self.gens = []
        for gid in range(first, first + local // self.g):
            gen = torch.Generator(device=dev)
            gen.manual_seed(int(seed) * 1000003 + gid)
            self.gens.append(gen)
        self.dev = dev

    def draw(self, kind, m):
        import torch
        f = torch.rand if kind == "uniform" else torch.randn
        return torch.cat([f((self.g, m), dtype=torch.float64, device=self.dev, generator=gen) for gen in self.gens], dim=0)


def make_ctcrw_device(n_tracks, n_steps, seed=20260103, device=0, k=10, sigma_obs=0.1,
                      irregular=True, rank=0, world=1, sim_tracks=None, dist_reduce=None,
                      shard_flags=0, time_shard=False, dist_gather=None, alias=None):
    """Build the rows of `n_tracks` tracks x `n_steps` steps on `device` (this rank's shard of
    `world * n_tracks` tracks; the data set does not depend on `world`, see GroupedRNG).

    sim_tracks: the single-long-track configuration -- the track (n_steps rows on this rank) is
    simulated as `sim_tracks` segments per rank that are stitched together (times and positions
    made continuous, across ranks too when time_shard is set).
    dist_reduce(tensor, op) -> tensor: all-reduce across ranks so that every shard uses the same
    knots and sum-to-zero constraint.  dist_gather(tensor) -> [world, numel] tensor (time_shard).
    time_shard: (with sim_tracks) this rank holds slab `rank` of `world` of ONE track of
    world * n_steps rows: its times follow the previous slab's, only rank 0 has the track start,
    only the last rank the track end (SSDE_SHARD_CONT_PREV / CONT_NEXT are set accordingly).
    alias: store the value slots that tau and nu (and mu1, mu2) share once per row (design.cuh
    alias flags: 22 column slots, 11 value slots); default on, SSDE_NO_ALIAS=1 gives the plain layout.
    Returns (engine, par, info)."""
    import os
    import torch
    if alias is None:
        alias = not os.environ.get("SSDE_NO_ALIAS")
    dev = torch.device("cuda", device)
    nd = 2
    T, m = n_tracks, n_steps
    single = sim_tracks is not None
    if single:
        assert n_tracks == 1
        T, m = sim_tracks, n_steps // sim_tracks
        assert T * m == n_steps
    n = T * m
    nslab = world if (time_shard or not single) else 1
    rng = GroupedRNG(seed, T * nslab, T, rank if nslab > 1 else 0, dev)
    if irregular:
        inc = 0.2 + 1.8 * rng.draw("uniform", m)
        inc[:, 0] = 0.0
        times = torch.cumsum(inc, dim=1)
        del inc
    else:
        times = torch.arange(m, dtype=torch.float64, device=dev).repeat(T, 1)

    def seg_offsets(x):
        """exclusive prefix sum of the per-segment quantity x over ALL segments of the track, in
        global segment order; the cumsum runs over the same full vector on every rank, so the
        offsets are bit-identical for every world size"""
        if time_shard and world > 1:
            allx = dist_gather(x.contiguous()).reshape(-1)
            o = torch.cumsum(allx, 0) - allx
            nxt.append(o[(rank + 1) * T] if rank < world - 1 else None)
            return o[rank * T:(rank + 1) * T]
        return torch.cumsum(x, 0) - x

    nxt = []            # [0]: time at which the next slab starts (time shards)

    if single:      # consecutive segments of one track: shift times so they increase overall
        times = times + seg_offsets(times[:, -1] + 1.0)[:, None]
        tmin, tmax = times.min().reshape(1), times.max().reshape(1)
        if time_shard and dist_reduce is not None:
            tmin, tmax = dist_reduce(tmin, "min"), dist_reduce(tmax, "max")
        s = (times - tmin) / (tmax - tmin)
    else:
        s = times / times[:, -1:]
    tau = torch.exp(0.5 * torch.sin(2 * math.pi * s))
    nu = torch.exp(0.3 * torch.cos(2 * math.pi * s))
    del s
    obs = torch.empty((n, nd), dtype=torch.float64, device=dev)
    for d in range(nd):
        e1, e2 = rng.draw("normal", m), rng.draw("normal", m)
        z = simulate_ctcrw_device(times.contiguous(), tau.contiguous(), nu.contiguous(), e1, e2, device)
        del e1, e2
        if single:
            z = z + seg_offsets(z[:, -1].clone())[:, None]
        z = z + sigma_obs * rng.draw("normal", m)
        obs[:, d] = z.reshape(-1)
        del z
    del tau, nu
    tflat = times.reshape(-1)
    # dt_i = t_{i+1} - t_i; 1 on the last row of each track
    dt = torch.ones(n, dtype=torch.float64, device=dev)
    dt[:-1] = tflat[1:] - tflat[:-1]
    flags = torch.full((n,), 4, dtype=torch.uint8, device=dev)
    if single and time_shard:
        first_slab, last_slab = rank == 0, rank == world - 1
        track_starts = np.array([0] if first_slab else [], dtype=np.int64)
        if first_slab:
            flags[0] |= 1
        if last_slab:
            flags[-1] |= 2
            dt[-1] = 1.0
        else:
            dt[-1] = nxt[0] - tflat[-1]          # first time of the next slab (= what the unsharded track has there)
        n_id = 1 if first_slab else 0
        shard_flags |= (0 if first_slab else L.SHARD_CONT_PREV) | (0 if last_slab else L.SHARD_CONT_NEXT)
    elif single:
        track_starts = np.array([0], dtype=np.int64)
        flags[0] |= 1
        flags[-1] |= 2
        dt[-1] = 1.0
        n_id = 1
    else:
        first = torch.arange(0, n, m, device=dev)
        flags[first] |= 1
        flags[first + (m - 1)] |= 2
        dt[first + (m - 1)] = 1.0
        track_starts = np.arange(0, n, m, dtype=np.int64)
        n_id = T
    # CTCRW never uses the dt of a track-start row: it carries the track index (row of a0)
    a0 = np.zeros((max(n_id, 1), 2 * nd))
    if n_id:
        dt[torch.as_tensor(track_starts, device=dev)] = torch.arange(n_id, dtype=torch.float64, device=dev)
        first_obs = obs[torch.as_tensor(track_starts, device=dev)].cpu().numpy()
        for d in range(nd):
            a0[:n_id, 2 * d] = first_obs[:, d]

    # spline design: knots over the global time range, sum-to-zero constraint from global means
    lo, hi = tflat.min().reshape(1), tflat.max().reshape(1)
    if dist_reduce is not None:
        lo, hi = dist_reduce(lo, "min"), dist_reduce(hi, "max")
    lo, hi = float(lo), float(hi)
    colsum = torch.zeros(k + 1, dtype=torch.float64, device=dev)
    CH = 1 << 22
    for c0 in range(0, n, CH):
        B = _bspline_torch(tflat[c0:c0 + CH], k, lo, hi)
        colsum[:k] += B.sum(0)
        colsum[k] += B.shape[0]
        del B
    if dist_reduce is not None:
        colsum = dist_reduce(colsum, "sum")
    means = (colsum[:k] / colsum[k]).cpu().numpy()
    Zc = _design.sum_to_zero_transform(means)
    S1 = Zc.T @ _design.second_diff_penalty(k) @ Zc + 1e-2 * np.eye(k - 1)
    S1 = 0.5 * (S1 + S1.T)
    Zt = torch.as_tensor(Zc, device=dev)
    nnz_row = 2 + 2 * k            # 2 mu intercepts + 2 x (intercept + k-1 spline columns)
    info4 = (C.c_int32 * 4)()
    L.load().ssde_layout_info(info4)
    lc, wt = int(info4[0]), int(info4[1])
    n_pad = int(L.load().ssde_padded_rows(n))
    nq = n_pad // wt
    # values: [n, value slots] natural -> [q][k][slot][lane].  The two mu intercepts hold the same
    # numbers, and so do the tau and nu blocks (same smooth of time): with `alias` a row stores
    # 1 + k values and the descriptor flags say that mu2 reads mu1's slot and nu reads tau's.
    nval_row = 1 + k if alias else nnz_row
    o_tau = 1 if alias else 2
    val = torch.zeros((n_pad, nval_row), dtype=torch.float64, device=dev)
    val[:n, 0] = 1.0
    val[:n, o_tau] = 1.0
    if not alias:
        val[:n, 1] = 1.0
        val[:n, 2 + k] = 1.0
    for c0 in range(0, n, CH):
        Bz = _bspline_torch(tflat[c0:c0 + CH], k, lo, hi) @ Zt
        c1 = min(c0 + CH, n)
        val[c0:c1, o_tau + 1:o_tau + k] = Bz
        if not alias:
            val[c0:c1, 3 + k:2 + 2 * k] = Bz
        del Bz
    val = val.reshape(nq, 32, lc, nval_row).permute(0, 2, 3, 1).contiguous().reshape(-1)
    p_fe, p_re = nd + 2, 2 * (k - 1)
    # theta = [coeff_fe (mu1, mu2, tau, nu intercepts) | coeff_re (tau spline, nu spline)];
    # every warp-tile uses the same 22 columns -> one shared column list
    pat = [0, 1, 2] + [p_fe + j for j in range(k - 1)] + [3] + [p_fe + (k - 1) + j for j in range(k - 1)]
    col = torch.as_tensor(np.asarray(pat, dtype=np.int32), device=dev)
    kmax = 1 | (1 << 8) | (k << 16) | (k << 24)
    desc = torch.empty((nq, 3), dtype=torch.int64, device=dev)
    desc[:, 0] = torch.arange(nq, dtype=torch.int64, device=dev) * (wt * nval_row)
    desc[:, 1] = 0
    wt_flags = 1 | (((1 << 10) | (3 << 14)) if alias else 0)      # WT_UNIFORM | mu2 -> mu1 | nu -> tau
    desc[:, 2] = kmax | (wt_flags << 32)
    obs_p = torch.stack([permute_rows(obs[:, d].contiguous(), n_pad, lc, 0.0) for d in range(nd)]).contiguous()
    dt_p = permute_rows(dt, n_pad, lc, 1.0)
    flags_p = permute_rows(flags, n_pad, lc, 255)

    keep = [desc, col, val, obs_p, dt_p, flags_p]
    pd = L.PackedDesc()
    pd.model, pd.n_dim, pd.n_par = L.SSDE_CTCRW, nd, nd + 2
    pd.n, pd.n_pad, pd.nnz = n, n_pad, n * nnz_row
    pd.d_desc, pd.d_col, pd.d_val = desc.data_ptr(), col.data_ptr(), val.data_ptr()
    pd.d_obs, pd.d_dt, pd.d_flags = obs_p.data_ptr(), dt_p.data_ptr(), flags_p.data_ptr()
    pd.p_fe, pd.p_re = p_fe, p_re
    S = sp.block_diag([S1, S1], format="csr")
    pd.S = _as_triplet(S, keep)
    ncol_re = np.array([k - 1, k - 1], dtype=np.int32)
    keep.append(ncol_re)
    pd.n_smooth = 2
    pd.ncol_re = ncol_re.ctypes.data_as(L.c_int32_p)
    pd.include_penalty = 1
    pd.n_ID = n_id
    ts = np.ascontiguousarray(track_starts if n_id else np.zeros(1, dtype=np.int64))
    a0c = np.ascontiguousarray(a0)
    keep += [ts, a0c]
    pd.track_starts = ts.ctypes.data_as(L.c_int64_p)
    pd.a0 = a0c.ctypes.data_as(L.c_double_p)
    pd.P0[0], pd.P0[1], pd.P0[2] = 1.0, 0.0, 10.0
    pd.device = device
    pd.shard_flags = shard_flags
    mu_cols = np.array([0, 1], dtype=np.int32)       # mu1.(Intercept), mu2.(Intercept)
    keep.append(mu_cols)
    pd.mu_cols = mu_cols.ctypes.data_as(L.c_int32_p)
    pd.n_mu_cols = 2
    torch.cuda.synchronize(dev)
    eng = Engine.from_packed(pd, keep)

    prng = np.random.default_rng(seed + 1)
    par = np.concatenate([[math.log(sigma_obs)], np.zeros(p_fe), np.zeros(2), 0.1 * prng.standard_normal(p_re)])
    info = {"n": n, "n_dim": nd, "nnz": n * nnz_row, "p_fe": p_fe, "p_re": p_re, "n_s": 2,
            "n_par": nd + 2, "n_tracks": n_id, "n_pad": n_pad, "tensors": dict(obs=obs, dt=dt, flags=flags, times=tflat),
            "S": S, "a0": a0, "track_starts": track_starts, "knots": (lo, hi), "Zc": Zc,
            "packed": dict(desc=desc, col=col, val=val, obs_p=obs_p, dt_p=dt_p, flags_p=flags_p, lc=lc, wt=wt, nnz_row=nnz_row,
                           nval_row=nval_row)}
    return eng, par, info


def slab_views(info, eng, cuts, device=0):
    """Engines over contiguous row ranges [cuts[r], cuts[r+1]) of a problem built by
    make_ctcrw_device, WITHOUT copying: every slab handle points into the same device arrays
    (ssde_create_packed).  Cuts must be multiples of the padding unit (1024 rows).  A cut inside a
    track gives time shards (SSDE_SHARD_CONT_PREV / CONT_NEXT), a cut at a track boundary gives
    track shards; the penalty stays with slab 0.  Used to check at BASELINE.json's full sizes that
    N shards reproduce the single-handle result."""
    import torch
    t = info["packed"]
    n, n_pad = info["n"], info["n_pad"]
    lc, wt, nnz_row = t["lc"], t["wt"], t["nnz_row"]
    starts_all = np.asarray(info["track_starts"], dtype=np.int64)
    a0_all = np.asarray(info["a0"], dtype=np.float64)
    flags_nat = info["tensors"]["flags"]
    out = []
    for r in range(len(cuts) - 1):
        lo, hi = int(cuts[r]), int(cuts[r + 1])
        assert lo % 1024 == 0 and (hi % 1024 == 0 or hi == n), "slab boundaries must be multiples of 1024 rows"
        m = hi - lo
        m_pad = int(L.load().ssde_padded_rows(m))
        assert lo + m_pad <= n_pad
        keep = list(eng._keep)
        pd = L.PackedDesc()
        pd.model, pd.n_dim, pd.n_par = L.SSDE_CTCRW, info["n_dim"], info["n_par"]
        pd.n, pd.n_pad, pd.nnz = m, m_pad, m * nnz_row
        pd.d_desc = t["desc"].data_ptr() + (lo // wt) * 24
        pd.d_col, pd.d_val = t["col"].data_ptr(), t["val"].data_ptr()      # val_off in the descriptors is absolute
        # per-row arrays: obs planes are n_pad apart in the full problem, the engine expects them
        # m_pad apart -> gather the planes of the slab (a copy of 16 B/row; everything else is a view)
        obs_p = torch.stack([t["obs_p"][d, lo:lo + m_pad] for d in range(info["n_dim"])]).contiguous()
        # track index carried by the dt slot of track-start rows must be slab-local
        sel = (starts_all >= lo) & (starts_all < hi)
        ts = starts_all[sel] - lo
        dt_p = t["dt_p"][lo:lo + m_pad]
        if sel.any() and int(np.flatnonzero(sel)[0]) != 0:
            dt_p = dt_p.clone()
            nat = permute_index(torch.as_tensor(ts, device=dt_p.device), lc)
            dt_p[nat] = torch.arange(ts.size, dtype=torch.float64, device=dt_p.device)
        keep += [obs_p, dt_p]
        pd.d_obs, pd.d_dt = obs_p.data_ptr(), dt_p.data_ptr()
        pd.d_flags = t["flags_p"].data_ptr() + lo
        pd.p_fe, pd.p_re = info["p_fe"], info["p_re"]
        pd.S = _as_triplet(info["S"], keep)
        ncol_re = np.array([info["p_re"] // 2, info["p_re"] // 2], dtype=np.int32)
        keep.append(ncol_re)
        pd.n_smooth, pd.ncol_re, pd.include_penalty = 2, ncol_re.ctypes.data_as(L.c_int32_p), 1
        a0 = np.ascontiguousarray(a0_all[sel]) if sel.any() else np.zeros((1, a0_all.shape[1]))
        tsc = np.ascontiguousarray(ts) if sel.any() else np.zeros(1, dtype=np.int64)
        keep += [a0, tsc]
        pd.n_ID = int(sel.sum())
        pd.track_starts, pd.a0 = tsc.ctypes.data_as(L.c_int64_p), a0.ctypes.data_as(L.c_double_p)
        pd.P0[0], pd.P0[1], pd.P0[2] = 1.0, 0.0, 10.0
        pd.device = device
        f_lo = int(flags_nat[lo]) if lo < n else 1
        f_hi = int(flags_nat[hi - 1])
        pd.shard_flags = (0 if (f_lo & 1) else L.SHARD_CONT_PREV) | (0 if (f_hi & 2) else L.SHARD_CONT_NEXT) | (L.SHARD_NO_PENALTY if r > 0 else 0)
        mu_cols = np.array([0, 1], dtype=np.int32)
        keep.append(mu_cols)
        pd.mu_cols, pd.n_mu_cols = mu_cols.ctypes.data_as(L.c_int32_p), 2
        out.append(Engine.from_packed(pd, keep))
    return out


def permute_index(rows, lc):
    """natural row index -> position in the warp-tile order (see permute_rows)."""
    wt = 32 * lc
    q = rows // wt
    r = rows - q * wt
    return q * wt + (r % lc) * 32 + r // lc


def alg_bytes_per_obs(n_dim, n_par, nnz_per_obs):
    """B_alg/n = (8 d + 8 + 4) + 2 [12 nnz/n + 4*2*n_par]."""
    return (8 * n_dim + 8 + 4) + 2 * (12 * nnz_per_obs + 8 * n_par)


# ------------------------------------------------------------------------------------------------
# OU with a random intercept per track: BASELINE configs[1] (64 x 1e5) and the OU half of configs[4]
# (4096 x 2.5e4)
# ------------------------------------------------------------------------------------------------
def make_ou_device(n_tracks, n_steps, seed=20260102, device=0, k=10, rank=0, world=1, shard_flags=0, alias=None):
    """OU, d = 1: ``mu, tau ~ s(time, k) + s(ID, bs = "re")``, ``kappa ~ 1`` for this
    rank's `n_tracks` of `world * n_tracks` tracks x `n_steps` regular steps, built on the device in the
    packed layout.  Columns: coeff_fe = (mu, tau, kappa intercepts); coeff_re = [mu.s(time) (k-1),
    mu.s(ID) (all tracks), tau.s(time), tau.s(ID)] as SDE$make_mat orders them (R/sde.R:412-421);
    S = blockdiag(S_time, I, S_time, I).  Each row has 23 nonzeros; a warp-tile that straddles two
    tracks carries both tracks' random-intercept columns (25 slots).  `alias` (default on, SSDE_NO_ALIAS=1
    turns it off): the mu and tau blocks hold the same numbers, so a row stores them once (design.cuh alias
    flags: 23 / 25 column slots, 12 / 13 value slots).  Returns (engine, par, info)."""
    import os
    import torch
    if alias is None:
        alias = not os.environ.get("SSDE_NO_ALIAS")
    dev = torch.device("cuda", device)
    T, m = int(n_tracks), int(n_steps)
    Ttot = T * world
    g0 = rank * T                                   # global index of this rank's first track
    n = T * m
    assert m >= 256, "tracks shorter than a warp-tile are not handled by this generator"
    info4 = (C.c_int32 * 4)()
    L.load().ssde_layout_info(info4)
    lc, wt = int(info4[0]), int(info4[1])
    n_pad = int(L.load().ssde_padded_rows(n))
    nq = n_pad // wt
    km1 = k - 1
    p_fe, p_re = 3, 2 * (km1 + Ttot)
    c_mu_spl, c_mu_re = p_fe, p_fe + km1
    c_tau_spl, c_tau_re = p_fe + km1 + Ttot, p_fe + 2 * km1 + Ttot

    # --- truth and data -------------------------------------------------------------------------
    gen_all = torch.Generator(device=dev)
    gen_all.manual_seed(int(seed) * 7919 + 1)
    u_all = 0.3 * torch.randn(Ttot, dtype=torch.float64, device=dev, generator=gen_all)     # random intercepts (all ranks draw the same)
    v_all = 0.2 * torch.randn(Ttot, dtype=torch.float64, device=dev, generator=gen_all)
    rng = GroupedRNG(seed, Ttot, T, rank, dev)
    t1 = torch.arange(m, dtype=torch.float64, device=dev)
    s = (t1 / (m - 1))[None, :]
    mu = 2.0 * torch.sin(2 * math.pi * s) + u_all[g0:g0 + T, None]
    tau = torch.exp(0.5 * torch.sin(2 * math.pi * s) + v_all[g0:g0 + T, None])
    kappa = torch.full((T, m), 1.5, dtype=torch.float64, device=dev)
    times = t1.repeat(T, 1).contiguous()
    e = rng.draw("normal", m)
    z = torch.zeros((T, m), dtype=torch.float64, device=dev)
    z[:, 0] = mu[:, 0]
    lib = L.load()
    torch.cuda.synchronize(dev)
    rc = lib.ssde_simulate_ou(device, T, m, times.data_ptr(), mu.contiguous().data_ptr(), tau.contiguous().data_ptr(),
                              kappa.data_ptr(), e.data_ptr(), z.data_ptr(), None)
    if rc != 0:
        raise L.EngineError(rc, lib.ssde_create_error().decode())
    torch.cuda.synchronize(dev)
    del e, mu, tau, kappa
    obs = z.reshape(-1)
    dt = torch.ones(n, dtype=torch.float64, device=dev)
    flags = torch.zeros(n, dtype=torch.uint8, device=dev)
    first = torch.arange(0, n, m, device=dev)
    flags[first] |= 1
    flags[first + (m - 1)] |= 2

    # --- spline block: identical for every track (regular times), sum-to-zero over one track ------
    lo, hi = 0.0, float(m - 1)
    B1 = _bspline_torch(t1, k, lo, hi)
    means = B1.mean(0).cpu().numpy()
    Zc = _design.sum_to_zero_transform(means)
    S1 = Zc.T @ _design.second_diff_penalty(k) @ Zc + 1e-2 * np.eye(km1)
    S1 = 0.5 * (S1 + S1.T)
    Bz1 = B1 @ torch.as_tensor(Zc, device=dev)                  # [m, k-1]

    # --- warp-tile descriptors ---------------------------------------------------------------------
    q = torch.arange(nq, device=dev, dtype=torch.int64)
    r_first = q * wt
    live = r_first < n
    r_last = torch.clamp(r_first + wt - 1, max=n - 1)
    tr_a = torch.where(live, r_first // m, torch.zeros_like(q))
    tr_b = torch.where(live, r_last // m, torch.zeros_like(q))
    two = live & (tr_b > tr_a)
    nre = torch.where(two, 2, 1)
    S_q = torch.where(live, 2 * (1 + km1) + 2 * nre + 1, torch.zeros_like(q))               # column slots
    SV_q = torch.where(live, (1 + km1) + nre + 1, torch.zeros_like(q)) if alias else S_q     # value slots
    val_off = torch.cumsum(SV_q * wt, 0) - SV_q * wt
    col_off = torch.cumsum(S_q, 0) - S_q
    n_val, n_col = int((SV_q * wt).sum()), int(S_q.sum())
    kp = (1 + km1) + nre
    kmax = torch.where(live, kp | (kp << 8) | (1 << 16), torch.zeros_like(q))
    wt_flags = 1 | ((1 << 10) if alias else 0)                  # WT_UNIFORM | tau's value slots = mu's
    desc = torch.empty((nq, 3), dtype=torch.int64, device=dev)
    desc[:, 0] = val_off
    desc[:, 1] = col_off
    desc[:, 2] = kmax | (wt_flags << 32)
    col = torch.zeros(max(n_col, 1), dtype=torch.int32, device=dev)
    spl = torch.arange(km1, device=dev, dtype=torch.int64)
    for is_two in (False, True):
        sel = torch.nonzero(live & (two == is_two)).reshape(-1)
        if sel.numel() == 0:
            continue
        ga, gb = g0 + tr_a[sel], g0 + tr_b[sel]
        ns = sel.numel()
        one = torch.ones(ns, dtype=torch.int64, device=dev)
        re_mu = [c_mu_re + ga] + ([c_mu_re + gb] if is_two else [])
        re_tau = [c_tau_re + ga] + ([c_tau_re + gb] if is_two else [])
        cols = torch.stack([0 * one] + [(c_mu_spl + j) * one for j in range(km1)] + re_mu
                           + [1 * one] + [(c_tau_spl + j) * one for j in range(km1)] + re_tau + [2 * one], dim=1)
        dest = col_off[sel][:, None] + torch.arange(cols.shape[1], device=dev)[None, :]
        col[dest.reshape(-1)] = cols.reshape(-1).to(torch.int32)
    del spl

    # --- values, in chunks of warp-tiles ---------------------------------------------------------------
    val = torch.zeros(max(n_val, 1), dtype=torch.float64, device=dev)
    CHQ = 1 << 12
    for q0 in range(0, nq, CHQ):
        q1 = min(q0 + CHQ, nq)
        rows = torch.arange(q0 * wt, q1 * wt, device=dev, dtype=torch.int64)
        real = rows < n
        rr = torch.clamp(rows, max=n - 1)
        trk = rr // m
        Bz = Bz1[rr - trk * m] * real[:, None]                   # [rows, k-1], zero on padding rows
        onev = real.to(torch.float64)
        for is_two in (False, True):
            sel = torch.nonzero((live & (two == is_two))[q0:q1]).reshape(-1)
            if sel.numel() == 0:
                continue
            n_blk = 1 if alias else 2                          # value blocks stored: mu (= tau) or mu, tau
            S_here = n_blk * ((1 + km1) + (2 if is_two else 1)) + 1
            rsel = (sel[:, None] * wt + torch.arange(wt, device=dev)[None, :]).reshape(-1)      # rows of the selected tiles
            V = torch.zeros((rsel.numel(), S_here), dtype=torch.float64, device=dev)
            o = 0
            for _p in range(n_blk):                            # mu block(, tau block)
                V[:, o] = onev[rsel]
                V[:, o + 1:o + 1 + km1] = Bz[rsel]
                o += 1 + km1
                if is_two:
                    in_a = (trk[rsel] == tr_a[q0:q1][sel].repeat_interleave(wt))
                    V[:, o] = onev[rsel] * in_a
                    V[:, o + 1] = onev[rsel] * (~in_a)
                    o += 2
                else:
                    V[:, o] = onev[rsel]
                    o += 1
            V[:, o] = onev[rsel]                               # kappa intercept
            blocks = V.reshape(sel.numel(), 32, lc, S_here).permute(0, 2, 3, 1).reshape(sel.numel(), -1)
            dest = val_off[q0:q1][sel][:, None] + torch.arange(S_here * wt, device=dev)[None, :]
            val[dest.reshape(-1)] = blocks.reshape(-1)
            del V, blocks, dest
    obs_p = permute_rows(obs.contiguous(), n_pad, lc, 0.0).reshape(1, n_pad).contiguous()
    dt_p = permute_rows(dt, n_pad, lc, 1.0)
    flags_p = permute_rows(flags, n_pad, lc, 255)

    keep = [desc, col, val, obs_p, dt_p, flags_p]
    pd = L.PackedDesc()
    pd.model, pd.n_dim, pd.n_par = L.SSDE_OU, 1, 3
    nnz = n * 23
    pd.n, pd.n_pad, pd.nnz = n, n_pad, nnz
    pd.d_desc, pd.d_col, pd.d_val = desc.data_ptr(), col.data_ptr(), val.data_ptr()
    pd.d_obs, pd.d_dt, pd.d_flags = obs_p.data_ptr(), dt_p.data_ptr(), flags_p.data_ptr()
    pd.p_fe, pd.p_re = p_fe, p_re
    S = sp.block_diag([S1, sp.identity(Ttot), S1, sp.identity(Ttot)], format="csr")
    pd.S = _as_triplet(S, keep)
    ncol_re = np.array([km1, Ttot, km1, Ttot], dtype=np.int32)
    keep.append(ncol_re)
    pd.n_smooth = 4
    pd.ncol_re = ncol_re.ctypes.data_as(L.c_int32_p)
    pd.include_penalty = 1
    pd.n_ID = T
    ts = np.ascontiguousarray(np.arange(0, n, m, dtype=np.int64))
    keep.append(ts)
    pd.track_starts = ts.ctypes.data_as(L.c_int64_p)
    pd.a0 = None
    pd.device = device
    pd.shard_flags = shard_flags
    pd.mu_cols, pd.n_mu_cols = None, 0
    torch.cuda.synchronize(dev)
    eng = Engine.from_packed(pd, keep)

    prng = np.random.default_rng(seed + 1)
    par = np.concatenate([[0.0, 0.0, math.log(1.5)], np.zeros(4), 0.1 * prng.standard_normal(p_re)])
    info = {"n": n, "n_dim": 1, "nnz": nnz, "p_fe": p_fe, "p_re": p_re, "n_s": 4, "n_par": 3, "n_tracks": T,
            "n_tracks_total": Ttot, "first_track": g0, "n_pad": n_pad, "S": S, "Bz1": Bz1, "m": m, "k": k,
            "stored_bytes_per_obs": (8.0 * n_val + 17.0 * n) / n,      # design values + dt + obs + flag
            "tensors": dict(obs=obs, dt=dt, flags=flags, times=times.reshape(-1))}
    return eng, par, info
