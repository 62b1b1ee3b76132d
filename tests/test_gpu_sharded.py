"""GPU tests of the sharded paths: a single track cut along time (three-stage scan protocol of
ssde_eval_stage) driven from one process, track shards, and both with one process per shard over
torch.distributed (NCCL with >= 2 GPUs, gloo with both ranks on the only GPU otherwise).  Reference numbers: the C oracle on the whole
problem."""
import os
import socket
import sys

import numpy as np
import pytest

from oracle import oracle_c
from smoothsde_b200 import sharded as S
from smoothsde_b200 import synth
from smoothsde_b200.engine import Engine

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))

NLLK_RTOL = 1e-10
GRAD_RTOL = 1e-7


def grad_err(g, g_ref):
    scale = np.maximum(np.abs(g_ref), 1e-3 * np.max(np.abs(g_ref)))
    return np.max(np.abs(g - g_ref) / scale)


def one_track(m, nd=2, miss=0.05, seed=9):
    dat, par, _ = synth.make_problem("CTCRW", 1, m, missing_frac=miss, n_dim=nd, seed=seed)
    par = par.copy()
    par[1:1 + nd] = [0.3, -0.2][:nd]
    return dat, par


@pytest.mark.parametrize("m,nd,nshard", [(5000, 2, 2), (5000, 2, 3), (2600, 1, 4), (7, 2, 2), (3000, 2, 8)])
def test_time_shards_on_one_gpu_match_oracle(m, nd, nshard):
    dat, par = one_track(m, nd)
    ref_v, ref_g = oracle_c.COracle(dat).eval(par, True)
    eng = S.TimeShardedEngine(dat, devices=[0] * nshard)
    v, g = eng.eval(par)
    assert abs(v - ref_v) <= NLLK_RTOL * max(abs(ref_v), 1.0), (v, ref_v)
    assert grad_err(g, ref_g) <= GRAD_RTOL
    v2, g2 = eng.eval(par)                       # epochs / incoming states are reset correctly
    assert abs(v2 - v) <= 1e-13 * abs(v) and grad_err(g2, g) <= 1e-12
    eng.close()


@pytest.mark.parametrize("model,nd", [("OU_SSM", 2), ("BM_SSM", 1)])
def test_time_shards_ssm_match_single_handle(model, nd):
    dat, par, _ = synth.make_problem(model, 1, 4000, missing_frac=0.05, n_dim=nd, seed=21)
    par = par.copy()
    par[1:1 + nd] = [0.3, -0.2][:nd]
    e1 = Engine.from_data(dat)
    ref_v, ref_g = e1.eval(par, 1)
    eng = S.TimeShardedEngine(dat, devices=[0, 0, 0])
    v, g = eng.eval(par)
    assert abs(v - ref_v) <= 1e-11 * max(abs(ref_v), 1.0), (v, ref_v)
    assert grad_err(g, ref_g) <= 1e-9
    eng.close(); e1.close()


def test_track_shards_solo_is_the_plain_engine():
    dat, par, _ = synth.make_problem("CTCRW", 6, 300, missing_frac=0.1, n_dim=2, seed=3)
    e1 = Engine.from_data(dat)
    e2 = S.TrackShardedEngine(dat, device=0)
    v1, g1 = e1.eval(par, 1)
    v2, g2 = e2.eval(par, 1)
    # same kernels on the same rows; the transposed design product is accumulated with atomics, so
    # the gradients agree to rounding, not bitwise
    assert v1 == v2 and np.max(np.abs(g1 - g2)) <= 1e-12 * np.max(np.abs(g1))
    d = np.linspace(-1, 1, par.size)
    _, _, h1 = e1.hvp(par, d)
    _, _, h2 = e2.hvp(par, d)
    assert np.max(np.abs(h1 - h2)) <= 1e-12 * np.max(np.abs(h1))
    e1.close(); e2.close()


# ---------------------------------------------------------------------------------------------
# one process per GPU over NCCL (needs >= 2 GPUs: `gpurun --gpus 2`)
# ---------------------------------------------------------------------------------------------
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _nccl_worker(rank, world, port, q, one_gpu):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dev = 0 if one_gpu else rank
    torch.cuda.set_device(dev)
    if one_gpu:          # NCCL refuses two ranks on one device: same processes and protocol, collectives over gloo
        dist.init_process_group("gloo", rank=rank, world_size=world)
    else:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    rank_dev = rank
    rank = dev
    try:
        comm = S.DistComm()
        dat, par, _ = synth.make_problem("CTCRW", 7, 400, missing_frac=0.1, n_dim=2, seed=3)
        e = S.TrackShardedEngine(dat, comm=comm, device=rank)
        v, g = e.eval(par, 1)
        _, _, hv = e.hvp(par, np.linspace(-1, 1, par.size))
        e.close()
        dat1, par1 = one_track(6000, 2)
        t = S.TimeShardedEngine(dat1, comm=comm, device=rank)
        tv, tg = t.eval(par1)
        t.close()
        q.put((rank_dev, v, g, hv, tv, tg))
    finally:
        dist.barrier()
        dist.destroy_process_group()


def test_one_process_per_shard_track_and_time_shards_match_oracle():
    """One process per shard with torch.distributed: over NCCL with one GPU per rank when the box has
    at least two GPUs; on a one-GPU box the same two-process protocol runs with both ranks on cuda:0
    and the collectives over gloo (never skipped)."""
    import torch
    ngpu = torch.cuda.device_count()
    one_gpu = ngpu < 2
    world = 2 if one_gpu else min(ngpu, 4)
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, world, port, q, one_gpu)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    dat, par, _ = synth.make_problem("CTCRW", 7, 400, missing_frac=0.1, n_dim=2, seed=3)
    co = oracle_c.COracle(dat)
    ref_v, ref_g = co.eval(par, True)
    d = np.linspace(-1, 1, par.size)
    k = 1e-3
    d1 = (co.eval(par + k * d)[1] - co.eval(par - k * d)[1]) / (2 * k)
    d2 = (co.eval(par + 0.5 * k * d)[1] - co.eval(par - 0.5 * k * d)[1]) / k
    ref_hv = (4 * d2 - d1) / 3
    dat1, par1 = one_track(6000, 2)
    ref_tv, ref_tg = oracle_c.COracle(dat1).eval(par1, True)
    for _, v, g, hv, tv, tg in res:
        assert abs(v - ref_v) <= NLLK_RTOL * abs(ref_v)
        assert grad_err(g, ref_g) <= GRAD_RTOL
        assert grad_err(hv, ref_hv) <= 1e-6
        assert abs(tv - ref_tv) <= NLLK_RTOL * abs(ref_tv)
        assert grad_err(tg, ref_tg) <= GRAD_RTOL


def test_time_shards_with_long_observation_gaps_take_the_full_summary_path():
    """No observation on the last 6000 rows of the first slab: its tail is NOT a constant map, so the
    protocol must fall back to whole-shard summary passes (stages 3 / 4) and still be exact."""
    dat, par = one_track(16000, 2, miss=0.02, seed=4)
    obs = dat["obs"].copy()
    obs[2000:8000] = np.nan
    dat = dict(dat, obs=obs)
    ref_v, ref_g = oracle_c.COracle(dat).eval(par, True)
    eng = S.TimeShardedEngine(dat, devices=[0, 0])
    v, g = eng.eval(par)
    assert abs(v - ref_v) <= NLLK_RTOL * max(abs(ref_v), 1.0), (v, ref_v)
    assert grad_err(g, ref_g) <= GRAD_RTOL
    eng.close()
