"""CPU tests of the multi-GPU host layer (smoothsde_b200/sharded.py): partitioning of the data
list by track ID and along time, penalty ownership, and the all-reduce path with two real
processes over gloo.  The evaluator is the oracle-backed fake engine (tests/fake_engine.py)."""
import os
import socket
import sys

import numpy as np
import pytest

from fake_engine import OracleEngine, oracle_shard_factory
from oracle import oracle_c
from smoothsde_b200 import sharded as S
from smoothsde_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))


def ragged_problem(model="CTCRW", lens=(3, 70, 41, 1, 26, 2), nd=2, seed=4):
    dat, par, _ = synth.make_problem(model, 1, sum(lens), missing_frac=0.1, n_dim=nd, seed=seed)
    ID = np.repeat(np.arange(len(lens)), lens).astype(float)
    i0 = np.r_[0, np.cumsum(lens)[:-1]]
    obs = dat["obs"].copy()
    obs[i0] = np.nan_to_num(obs[i0])
    dat = dict(dat, ID=ID, obs=obs)
    if model == "CTCRW":
        a0 = np.zeros((len(lens), 2 * nd))
        for d in range(nd):
            a0[:, 2 * d] = obs[i0, d]
        dat["a0"] = a0
    return dat, par


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_split_tracks_keeps_tracks_whole_and_covers_all_rows(world):
    ID = np.repeat(np.arange(6), [3, 70, 41, 1, 26, 2])
    parts = S.split_tracks(ID, world)
    assert len(parts) == world and parts[0][0] == 0 and parts[-1][1] == ID.size
    for (lo, hi), (lo2, _) in zip(parts[:-1], parts[1:]):
        assert hi == lo2
    b = set(S.track_bounds(ID).tolist())
    for lo, hi in parts:
        assert lo in b and hi in b


def test_split_time_and_shard_flags():
    dat, par, _ = synth.make_problem("CTCRW", 1, 50, n_dim=2, seed=1)
    parts = S.split_time(50, 4)
    assert parts[0][0] == 0 and parts[-1][1] == 50
    flags = []
    for lo, hi in parts:
        sub, cp, cn, t_next = S.shard_rows(dat, lo, hi)
        flags.append((cp, cn))
        assert sub["X_re"].shape == (4 * (hi - lo), dat["X_re"].shape[1])
        assert sub["a0"].shape[0] == (1 if lo == 0 else 0)
        if cn:
            assert t_next == dat["times"][hi]
    assert flags == [(False, True), (True, True), (True, True), (True, False)]


@pytest.mark.parametrize("model,nd", [("CTCRW", 2), ("OU", 1), ("BM", 2)])
@pytest.mark.parametrize("world", [2, 3])
def test_track_shards_sum_to_the_full_objective(model, nd, world):
    dat, par = ragged_problem(model, nd=nd)
    ref_v, ref_g = oracle_c.COracle(dat).eval(par, True)
    tot_v, tot_g = 0.0, 0.0
    for r, (lo, hi) in enumerate(S.split_tracks(dat["ID"], world)):
        sub, cp, cn, _ = S.shard_rows(dat, lo, hi)
        assert not cp and not cn
        v, g = OracleEngine(sub, add_penalty=(r == 0)).eval(par, 1)
        tot_v, tot_g = tot_v + v, tot_g + g
    assert abs(tot_v - ref_v) <= 1e-12 * abs(ref_v)
    assert np.max(np.abs(tot_g - ref_g)) <= 1e-10 * max(1.0, np.max(np.abs(ref_g)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dat, par = ragged_problem("CTCRW", nd=2)
        eng = S.TrackShardedEngine(dat, comm=S.DistComm(), device=None, engine_factory=oracle_shard_factory)
        v, g = eng.eval(par, order=1)
        d = np.linspace(-1, 1, par.size)
        _, _, hv = eng.hvp(par, d)
        q.put((rank, eng.lo, eng.hi, v, g, hv))
    finally:
        dist.barrier()
        dist.destroy_process_group()


def test_two_process_gloo_all_reduce_matches_single_process():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    dat, par = ragged_problem("CTCRW", nd=2)
    full = OracleEngine(dat)
    ref_v, ref_g = full.eval(par, 1)
    _, _, ref_hv = full.hvp(par, np.linspace(-1, 1, par.size))
    assert res[0][2] == res[1][1] and res[0][1] == 0 and res[1][2] == dat["ID"].size      # disjoint cover
    for _, _, _, v, g, hv in res:                                                         # every rank has the sum
        assert abs(v - ref_v) <= 1e-12 * abs(ref_v)
        assert np.max(np.abs(g - ref_g)) <= 1e-10 * max(1.0, np.max(np.abs(ref_g)))
        assert np.max(np.abs(hv - ref_hv)) <= 1e-6 * max(1.0, np.max(np.abs(ref_hv)))


def test_shard_rows_carries_user_H_and_decay_times():
    """Per-row extras follow their rows: H_array[, , lo:hi] (nllk_ctcrw.hpp:203-205) and the
    matching entries of every parameter block of t_decay (nllk_sde.hpp:53)."""
    dat, par, info = synth.make_problem("CTCRW", 4, 30, n_dim=2, seed=2, k=5)
    n = info["n"]
    dat["H_array"] = np.arange(4 * n, dtype=float).reshape(2, 2, n)
    (lo0, hi0), (lo1, hi1) = S.split_tracks(dat["ID"], 2)
    sub, cp, cn, _ = S.shard_rows(dat, lo1, hi1)
    assert not cp and not cn and np.array_equal(sub["H_array"], dat["H_array"][:, :, lo1:hi1])
    assert sub["a0"].shape[0] == np.unique(np.asarray(dat["ID"])[lo1:hi1]).size
    dat2, par2, info2 = synth.make_problem("OU", 4, 30, n_dim=1, seed=2, k=5)
    n2, n_par = info2["n"], 3
    dat2["t_decay"] = np.arange(n_par * n2, dtype=float)
    dat2["col_decay"], dat2["ind_decay"] = np.array([1, 2]), np.array([1, 1])
    lo, hi = S.split_tracks(dat2["ID"], 2)[1]
    sub2, _, _, _ = S.shard_rows(dat2, lo, hi)
    want = np.concatenate([dat2["t_decay"][j * n2 + lo:j * n2 + hi] for j in range(n_par)])
    assert np.array_equal(sub2["t_decay"], want) and sub2["X_re"].shape[0] == n_par * (hi - lo)
    assert list(sub2["col_decay"]) == [1, 2]


def test_split_tracks_gives_every_rank_a_track_even_when_unbalanced():
    """track sizes [100, 1, 1, 1] on 4 ranks used to leave ranks empty (which then blocked the
    others in the first collective); with fewer tracks than ranks EVERY rank must raise."""
    from smoothsde_b200 import sharded as S

    def ids(sizes):
        return np.repeat(np.arange(len(sizes)), sizes)
    for sizes, world in (([100, 1, 1, 1], 4), ([1, 1, 1, 100], 4), ([50, 1, 1, 50, 3], 3), ([7] * 16, 8), ([5, 9, 2], 3)):
        groups = S.split_tracks(ids(sizes), world)
        assert len(groups) == world and groups[0][0] == 0 and groups[-1][1] == sum(sizes)
        assert all(hi > lo for lo, hi in groups), (sizes, groups)
        assert all(groups[r][1] == groups[r + 1][0] for r in range(world - 1))
        bounds = set(np.cumsum([0] + list(sizes)).tolist())
        assert all(lo in bounds and hi in bounds for lo, hi in groups)          # whole tracks only
    groups = S.split_tracks(ids([5, 5]), 4)
    assert sum(hi > lo for lo, hi in groups) == 2

    class FakeComm:
        def __init__(self, rank):
            self.rank, self.world = rank, 4
    dat = {"ID": ids([5, 5]).astype(float)}
    for rank in range(4):                        # the same error on every rank, before any collective
        with pytest.raises(ValueError, match="cannot be spread over 4 ranks"):
            S.TrackShardedEngine(dat, comm=FakeComm(rank), device=None, engine_factory=lambda *a: None)
