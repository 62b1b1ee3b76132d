"""Multi-GPU host layer: one process per GPU (``torch.distributed``, NCCL over NVLink), or several
shards driven by one process.

The reference has no parallelism at all: one sequential loop over every row of every track
(nllk_ctcrw.hpp:195-247, nllk_sde.hpp:77-84).  What makes the path shard is in the model:

* tracks are independent given the parameters (the filter is re-initialised at every ID change,
  nllk_ctcrw.hpp:196-200; nllk_sde.hpp:79 skips the transition across it), so ranks hold
  disjoint groups of whole tracks -- their rows of ``ID / times / obs`` and the matching rows of
  every parameter block of ``X_fe / X_re`` -- and the packed ``[nllk, gradient]`` (or a Hessian-
  vector product) is summed with ONE all-reduce per evaluation.  The smoothing penalty is added
  by rank 0 only (``SSDE_SHARD_NO_PENALTY`` elsewhere).  -> :class:`TrackShardedEngine`
* the filter and its adjoint along ONE track are associative scans (csrc/ctcrw_math.cuh), so a
  single long track is cut into contiguous time slabs; per evaluation every shard contributes one
  composite element per sweep (17 + 11 doubles for d = 2), exchanged with two all-gathers, then
  the same all-reduce.  -> :class:`TimeShardedEngine` (CTCRW)

Collectives run on a dedicated torch stream that is also handed to the C ABI, so kernels and
collectives are ordered on the device without host synchronisation.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from . import _lib as L


# --------------------------------------------------------------------------------------------
# partitioning of the reference's data list
# --------------------------------------------------------------------------------------------
def n_sde_par(dat):
    obs = np.asarray(dat["obs"])
    nd = 1 if obs.ndim == 1 else obs.shape[1]
    return nd + 1 if dat["type"] in ("BM", "BM_SSM") else nd + 2


def track_bounds(ID):
    """Row index of the first row of every track, plus n (tracks = runs of equal ID)."""
    ID = np.asarray(ID)
    starts = np.flatnonzero(np.r_[True, ID[1:] != ID[:-1]])
    return np.r_[starts, ID.size]


def split_tracks(ID, world):
    """Contiguous groups of whole tracks, balanced by row count: [(lo, hi)] * world.  With at least
    as many tracks as ranks every rank gets at least one track (cut r is the track boundary nearest
    to r/world of the rows among those that leave one track for every rank before and after it);
    with fewer tracks than ranks the trailing groups are empty -- every rank sees the same ID
    vector, so every rank can raise the same error (TrackShardedEngine)."""
    b = track_bounds(ID)
    n, nt = int(b[-1]), b.size - 1
    if nt < world:
        cuts = [int(b[min(r, nt)]) for r in range(world + 1)]
        return [(cuts[r], cuts[r + 1]) for r in range(world)]
    idx = [0]                                   # indices into b
    for r in range(1, world):
        lo, hi = idx[-1] + 1, nt - (world - r)  # leave >= 1 track for this group and for each later one
        target = n * r / world
        j = lo + int(np.argmin(np.abs(b[lo:hi + 1] - target)))
        idx.append(j)
    idx.append(nt)
    return [(int(b[idx[r]]), int(b[idx[r + 1]])) for r in range(world)]


def split_time(n, world):
    """Contiguous time slabs of one track: [(lo, hi)] * world."""
    cuts = [int(round(n * r / world)) for r in range(world + 1)]
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def shard_rows(dat, lo, hi):
    """Rows lo..hi-1 of the data list (R/sde.R:528-598): rows of ID / times / obs, the matching
    rows of every SDE-parameter block of X_fe / X_re, and a0 of the tracks that START in the
    range.  Returns (sub_dat, cont_prev, cont_next, t_next)."""
    ID = np.asarray(dat["ID"])
    n = ID.size
    n_par = n_sde_par(dat)
    obs = np.asarray(dat["obs"], dtype=float)
    if obs.ndim == 1:
        obs = obs[:, None]
    sub = {k: v for k, v in dat.items() if k not in ("ID", "times", "obs", "X_fe", "X_re", "a0", "H_array", "t_decay")}
    td = dat.get("t_decay")
    if td is not None and np.size(td) > 1:           # one entry per row of X_re: follows the rows of every block
        td = np.asarray(td, dtype=float)
        sub["t_decay"] = np.concatenate([td[j * n + lo:j * n + hi] for j in range(n_par)])
    sub["ID"], sub["times"], sub["obs"] = ID[lo:hi], np.asarray(dat["times"])[lo:hi], obs[lo:hi]
    for nm in ("X_fe", "X_re"):
        X = sp.csr_matrix(dat[nm])
        sub[nm] = sp.vstack([X[j * n + lo:j * n + hi] for j in range(n_par)], format="csr")
    cont_prev = lo > 0 and ID[lo - 1] == ID[lo]
    cont_next = hi < n and ID[hi] == ID[hi - 1]
    if dat["type"] in L.KALMAN_TYPES:
        b = track_bounds(ID)[:-1]
        a0 = np.asarray(dat["a0"], dtype=float).reshape(b.size, -1)
        sub["a0"] = a0[(b >= lo) & (b < hi)]
        H = dat.get("H_array")
        if H is not None and np.size(H) > 1:         # user measurement covariances follow their rows
            sub["H_array"] = np.asarray(H, dtype=float)[:, :, lo:hi]
    t_next = float(np.asarray(dat["times"])[hi]) if cont_next else 0.0
    return sub, bool(cont_prev), bool(cont_next), t_next


# --------------------------------------------------------------------------------------------
# communication: torch.distributed, or nothing (world = 1)
# --------------------------------------------------------------------------------------------
class DistComm:
    """Thin wrapper over a torch.distributed process group: NCCL on GPUs (device tensors go straight
    to the collective, ordered on the current stream); gloo in the CPU tests and when several ranks
    share ONE GPU (NCCL refuses two ranks on a device) -- device tensors are then staged through the
    host, which also orders them after the kernels of the current stream."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.staged = dist.get_backend(group) != "nccl"

    def all_reduce_sum(self, t):
        if self.world > 1:
            if self.staged and t.is_cuda:
                h = t.cpu()
                self.dist.all_reduce(h, op=self.dist.ReduceOp.SUM, group=self.group)
                t.copy_(h)
            else:
                self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t

    def all_gather(self, t):
        import torch
        out = torch.empty((self.world * t.numel(),), dtype=t.dtype, device=t.device)
        if self.world > 1:
            if self.staged:
                parts = [torch.empty(t.numel(), dtype=t.dtype) for _ in range(self.world)]
                self.dist.all_gather(parts, t.reshape(-1).cpu().contiguous(), group=self.group)
                out.copy_(torch.cat(parts))
            else:
                self.dist.all_gather_into_tensor(out, t.contiguous(), group=self.group)
        else:
            out.copy_(t.reshape(-1))
        return out


class SoloComm:
    rank, world = 0, 1

    def all_reduce_sum(self, t):
        return t

    def all_gather(self, t):
        return t.reshape(-1).clone()


def _default_factory(dat, device, shard_flags, t_next):
    from .engine import Engine
    return Engine.from_data(dat, device=device, shard_flags=shard_flags, t_next=t_next)


# --------------------------------------------------------------------------------------------
# many tracks: shard by track ID
# --------------------------------------------------------------------------------------------
class TrackShardedEngine:
    """This rank's group of tracks + one all-reduce per evaluation.  Same ``eval`` / ``hvp`` /
    ``hessian`` interface as :class:`smoothsde_b200.engine.Engine`; every rank gets the full result."""

    def __init__(self, dat, comm=None, device=None, engine_factory=None):
        import torch
        self.comm = comm if comm is not None else SoloComm()
        rank, world = self.comm.rank, self.comm.world
        groups = split_tracks(dat["ID"], world)
        if any(hi <= lo for lo, hi in groups):
            # raised on EVERY rank (all see the same ID vector): no rank is left waiting in a collective
            raise ValueError(f"{track_bounds(dat['ID']).size - 1} track(s) cannot be spread over {world} ranks: "
                             "use at most as many ranks as tracks")
        self.lo, self.hi = groups[rank]
        sub, cp, cn, _ = shard_rows(dat, self.lo, self.hi)
        assert not cp and not cn
        flags = L.SHARD_NO_PENALTY if rank > 0 else 0
        self.device = device
        factory = engine_factory or _default_factory
        self.engine = factory(sub, 0 if device is None else device, flags, 0.0)
        self._init_buffers(device)

    @classmethod
    def from_engine(cls, engine, comm, device):
        """Adopt this rank's already-built shard (e.g. from devgen, whose 1e8-row problems never
        exist as a host data list); rank > 0 must have been built with SSDE_SHARD_NO_PENALTY."""
        self = cls.__new__(cls)
        self.comm, self.device, self.engine = comm, device, engine
        self.lo = self.hi = None
        self._init_buffers(device)
        return self

    def _init_buffers(self, device):
        import torch
        self.n_par, self.layout = self.engine.n_par, self.engine.layout
        self.on_device = device is not None and hasattr(self.engine, "eval_device")
        if self.on_device:
            self.dev = torch.device("cuda", device)
            self.stream = torch.cuda.Stream(device=self.dev)
            self._par = torch.zeros(self.n_par, dtype=torch.float64, device=self.dev)
            self._dir = torch.zeros(self.n_par, dtype=torch.float64, device=self.dev)
            self._out = torch.zeros(2 * self.n_par + 2, dtype=torch.float64, device=self.dev)
            self._hpar = torch.zeros(self.n_par, dtype=torch.float64).pin_memory()
            self._hout = torch.zeros(2 * self.n_par + 2, dtype=torch.float64).pin_memory()

    # ---- host-buffer interface (what obj$fn / obj$gr see)
    def eval(self, par, order=1):
        import torch
        par = np.ascontiguousarray(par, dtype=np.float64)
        np_ = self.n_par
        if not self.on_device:
            v, g = self.engine.eval(par, order=order)
            buf = torch.zeros(np_ + 1, dtype=torch.float64)
            buf[0] = v
            if order >= 1:
                buf[1:] = torch.as_tensor(g)
            self.comm.all_reduce_sum(buf)
            out = buf.numpy()
            return float(out[0]), (out[1:].copy() if order >= 1 else None)
        with torch.cuda.stream(self.stream):
            self._hpar.copy_(torch.as_tensor(par))
            self._par.copy_(self._hpar, non_blocking=True)
            self.engine.eval_device(self._par.data_ptr(), self._out.data_ptr(), order, self.stream.cuda_stream)
            # [nllk, gradient, status]: the status word is summed too (non-zero on any rank = failure)
            self.comm.all_reduce_sum(self._out[:np_ + 2])
            self._hout[:np_ + 2].copy_(self._out[:np_ + 2], non_blocking=True)
        self.stream.synchronize()
        out = self._hout.numpy()
        if out[np_ + 1] != 0.0:
            raise L.EngineError(5, "device-side failure on some rank")
        return float(out[0]), (out[1:np_ + 1].copy() if order >= 1 else None)

    def hvp(self, par, dirs):
        """(nllk, grad, H @ dirs): one tangent pass + one all-reduce per direction."""
        import torch
        par = np.ascontiguousarray(par, dtype=np.float64)
        dirs = np.asarray(dirs, dtype=np.float64)
        one = dirs.ndim == 1
        D = dirs.reshape(self.n_par, -1)
        np_ = self.n_par
        hv = np.zeros_like(D)
        v, g = None, None
        if not self.on_device:
            for j in range(D.shape[1]):
                v, g, h = self.engine.hvp(par, D[:, j])
                buf = torch.as_tensor(np.concatenate([[v], g, h]))
                self.comm.all_reduce_sum(buf)
                out = buf.numpy()
                v, g, hv[:, j] = float(out[0]), out[1:np_ + 1].copy(), out[np_ + 1:]
            return v, g, (hv[:, 0] if one else hv)
        with torch.cuda.stream(self.stream):
            self._par.copy_(torch.as_tensor(par), non_blocking=False)
            for j in range(D.shape[1]):
                self._dir.copy_(torch.as_tensor(np.ascontiguousarray(D[:, j])), non_blocking=False)
                self.engine.hvp_device(self._par.data_ptr(), self._dir.data_ptr(), self._out.data_ptr(),
                                       self._out.data_ptr() + 8 * (np_ + 2), self.stream.cuda_stream)
                self.comm.all_reduce_sum(self._out)
                out = self._out.cpu().numpy()
                hv[:, j] = out[np_ + 2:]
        if out[np_ + 1] != 0.0:
            raise L.EngineError(5, "device-side failure on some rank")
        return float(out[0]), out[1:np_ + 1].copy(), (hv[:, 0] if one else hv)

    def hessian(self, par):
        v, g, H = self.hvp(par, np.eye(self.n_par))
        return v, g, 0.5 * (H + H.T)

    def close(self):
        self.engine.close()


# --------------------------------------------------------------------------------------------
# one long track: shard along time (CTCRW)
# --------------------------------------------------------------------------------------------
def _stage_driver(shards, par_tensors, gather, reduce_, streams, fallbacks=None):
    """The stages of ssde_eval_stage for a list of local shards (one per rank in the distributed
    case).  gather(list of per-shard tensors) -> list of gathered tensors, one per local shard;
    reduce_(list of out tensors) sums them in place across all shards.

    Summary passes first run over the TAIL of each shard only (stage 0 / the summary part of stage
    1): with observations the filter forgets, so the last few thousand rows of a shard almost
    always are its exact composite scan element.  Every element carries a constant-map flag; if
    any shard's flag is unset (long gaps without observations), the pass is repeated over whole
    shards (stages 3 / 4).  Reading the flags is the only host synchronisation of an evaluation."""
    import torch

    def run(stage, which, gathered=None):
        outs = []
        for i, ((eng, me, nsh), par, st) in enumerate(zip(shards, par_tensors, streams)):
            with torch.cuda.stream(st):
                t = torch.empty(eng.shard_elem_doubles(which) if which is not None else eng.n_par + 2,
                                dtype=torch.float64, device=par.device)
                g = gathered[i] if gathered is not None else None
                eng.eval_stage(stage, par.data_ptr(), t.data_ptr(), g.data_ptr() if g is not None else 0, nsh, me,
                               stream_ptr=st.cuda_stream)
            outs.append(t)
        return outs

    def all_const(g, which):
        # every rank reads the same gathered flags, so all ranks take the same branch; the copy is
        # ordered after the gather on the shard's stream
        k = shards[0][0].shard_elem_doubles(which)
        with torch.cuda.stream(streams[0]):
            flags = g[0][k - 1::k].cpu()
        return bool((flags > 0.5).all())

    g0 = gather(run(0, 0))
    if not all_const(g0, 0):
        g0 = gather(run(3, 0))
        if fallbacks is not None:
            fallbacks[0] += 1
    g1 = gather(run(1, 1, g0))
    if not all_const(g1, 1):
        g1 = gather(run(4, 1))
        if fallbacks is not None:
            fallbacks[1] += 1
    return reduce_(run(2, None, g1))


class TimeShardedEngine:
    """One CTCRW track cut along time.  Distributed use: one slab per rank (``comm`` = DistComm).
    Single-process use: ``devices`` = list of CUDA ordinals (repeats allowed), all slabs driven
    from this process, elements exchanged with device-to-device copies."""

    fallbacks = None        # set to [0, 0] to count whole-shard summary passes [forward stage 3, adjoint stage 4]

    def __init__(self, dat, comm=None, device=0, devices=None, engine_factory=None):
        import torch
        if dat["type"] not in L.KALMAN_TYPES:
            raise L.EngineError(3, "time sharding exists for the Kalman models only")
        if track_bounds(dat["ID"]).size != 2:
            raise ValueError("TimeShardedEngine takes ONE track; use TrackShardedEngine for many")
        n = np.asarray(dat["ID"]).size
        factory = engine_factory or _default_factory
        self.local = devices is not None
        if self.local:
            world, ranks, devs = len(devices), list(range(len(devices))), list(devices)
            self.comm = SoloComm()
        else:
            self.comm = comm if comm is not None else SoloComm()
            world, ranks, devs = self.comm.world, [self.comm.rank], [device]
        slabs = split_time(n, world)
        self.shards, self.streams, self.devs = [], [], []
        for r, dv in zip(ranks, devs):
            lo, hi = slabs[r]
            sub, cp, cn, t_next = shard_rows(dat, lo, hi)
            flags = (L.SHARD_CONT_PREV if cp else 0) | (L.SHARD_CONT_NEXT if cn else 0) | (L.SHARD_NO_PENALTY if r > 0 else 0)
            self.shards.append((factory(sub, dv, flags, t_next), r, world))
            self.devs.append(torch.device("cuda", dv))
            self.streams.append(torch.cuda.Stream(device=self.devs[-1]))
        self.n_par, self.layout = self.shards[0][0].n_par, self.shards[0][0].layout
        self.world = world

    @classmethod
    def from_engines(cls, engines, devices):
        """Adopt already-built shard engines (in time order), all driven from this process."""
        import torch
        self = cls.__new__(cls)
        self.local, self.comm, self.world = True, SoloComm(), len(engines)
        self.shards = [(e, r, len(engines)) for r, e in enumerate(engines)]
        self.devs = [torch.device("cuda", d) for d in devices]
        self.streams = [torch.cuda.Stream(device=d) for d in self.devs]
        self.n_par, self.layout = engines[0].n_par, engines[0].layout
        return self

    def _gather(self, ts):
        import torch
        if not self.local:
            with torch.cuda.stream(self.streams[0]):
                return [self.comm.all_gather(ts[0])]
        for st in self.streams:
            st.synchronize()
        return [torch.cat([t.to(dv) for t in ts]) for dv in self.devs]

    def _reduce(self, outs):
        import torch
        if not self.local:
            with torch.cuda.stream(self.streams[0]):
                self.comm.all_reduce_sum(outs[0])
                res = outs[0].cpu()
            return res.numpy()
        for st in self.streams:
            st.synchronize()
        return sum(o.cpu() for o in outs).numpy()

    @classmethod
    def from_engine_distributed(cls, engine, comm, device):
        """This rank's time slab (already built, e.g. by devgen) + a torch.distributed group."""
        import torch
        self = cls.__new__(cls)
        self.local, self.comm, self.world = False, comm, comm.world
        self.shards = [(engine, comm.rank, comm.world)]
        self.devs = [torch.device("cuda", device)]
        self.streams = [torch.cuda.Stream(device=self.devs[0])]
        self.n_par, self.layout = engine.n_par, engine.layout
        return self

    def eval_device(self, par_dev):
        """Distributed use, no host synchronisation: par_dev is a device tensor; returns the device
        tensor [nllk, gradient, status] (all-reduced), ordered on self.streams[0]."""
        import torch

        def reduce_(outs):
            with torch.cuda.stream(self.streams[0]):
                self.comm.all_reduce_sum(outs[0])
            return outs[0]
        return _stage_driver(self.shards, [par_dev], self._gather, reduce_, self.streams, self.fallbacks)

    def eval(self, par, order=1):
        """(nllk, grad); the stage protocol always computes the gradient."""
        import torch
        par = np.ascontiguousarray(par, dtype=np.float64)
        pts = []
        for dv, st in zip(self.devs, self.streams):
            with torch.cuda.stream(st):
                pts.append(torch.as_tensor(par).to(dv))
        out = _stage_driver(self.shards, pts, self._gather, self._reduce, self.streams, self.fallbacks)
        if out[self.n_par + 1] != 0.0:
            raise L.EngineError(5, "device-side failure on some shard")
        return float(out[0]), (out[1:self.n_par + 1].copy() if order >= 1 else None)

    def close(self):
        for eng, _, _ in self.shards:
            eng.close()
