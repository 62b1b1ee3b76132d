#!/usr/bin/env python
"""Benchmark of the hot path: joint nllk + gradient evaluations of a 1e8-observation CTCRW model.

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference          # CPU arm: the reference's own templates (oracle/_ref)

A "step" is one evaluation of the joint penalised nllk and its gradient (what obj$fn + obj$gr do
in SDE$fit(), R/sde.R:694-697) over the whole synthetic data set, which is resident in HBM.
Headline workload (BASELINE.json configs[2], the CTCRW 1e8-observation case the metric is quoted on):
1024 simulated tracks x 1e5 irregular steps, d = 2, tau, nu ~ s(time, k = 10), mu fixed at 0.
With N > 1 GPUs the tracks are split across ranks (strong scaling: the total stays 1.024e8
rows) and the packed [nllk, gradient] vector is summed with one NCCL all-reduce per evaluation.
The data set is the SAME at every N (devgen.GroupedRNG) and the line's `parity_vs_n1` compares nllk /
gradient with the stored 1-GPU result (profiles/bench_nllk_n1.json; `--write-n1` refreshes it).

The same JSON line also carries (default `--workload both`):
  single    BASELINE configs[3]: ONE track of 1.024e8 rows, sharded along time (2 all-gathers of a scan
            element + 1 all-reduce per evaluation), with the count of whole-shard fallback passes;
  configs   the BM / OU path: configs[4]'s OU half (4096 tracks x 25 000 steps, mu, tau ~ s(time) +
            s(ID, bs = "re"), p_re = 8210) and configs[1] (64 x 1e5), tracks sharded over the ranks;
  roofline  frac (algorithmic bytes of SURVEY 8(d) / launch time / peak) AND dram_frac (DRAM bytes the ncu
            capture of the same kernel measured, scaled by rows / launch time / peak);
  cpu_baseline (N = 1 only)  the reference arm on a bounded sample, all host cores and one thread, with the
            C port beside it.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "joint nllk+grad obs/s (CTCRW Kalman, 1e8 obs) at 1/2/4/8 B200 vs TMB CPU"
UNIT = "obs*eval/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--tracks", type=int, default=1024)
    ap.add_argument("--track-steps", type=int, default=100000)
    ap.add_argument("--cpu-tracks", type=int, default=128)
    ap.add_argument("--cpu-track-steps", type=int, default=8192)
    ap.add_argument("--cpu-b1-tracks", type=int, default=8, help="tracks of the CPU sample timed on ONE thread (B1)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-configs", action="store_true", help="skip the OU configurations reported under `configs`")
    ap.add_argument("--workload", default="both", choices=["both", "tracks", "single"],
                    help="tracks: BASELINE configs[2] (the headline line); single: configs[3], ONE track of "
                         "tracks*track_steps rows, sharded along time over the ranks; both (default): the headline "
                         "line carries the time-sharded measurement under the key `single`")
    ap.add_argument("--write-n1", action="store_true",
                    help="(1 GPU) store this run's nllk as the N = 1 value that multi-GPU runs are checked against")
    return ap.parse_args()


def host_cores():
    """Cores this process may run on -- NOT omp_get_max_threads(): torch.distributed.run exports
    OMP_NUM_THREADS=1, which must not shrink the CPU baseline."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ------------------------------------------------------------------------------------------------
# CPU arm.  kind "reference": the reference's own objective (/root/reference/src/smoothSDE.cpp +
# src/nllk/*.hpp, unmodified, compiled against oracle/tmb_shim/TMB.hpp into oracle/_ref), value +
# gradient by one reverse sweep of the AD tape per track, tracks spread over a thread pool.
# kind "port" (reported beside it): oracle/oracle_c.c, our C restatement with a hand-written adjoint.
# ------------------------------------------------------------------------------------------------
def cpu_problem(args):
    from smoothsde_b200 import synth
    dat, par, info = synth.make_problem("CTCRW", args.cpu_tracks, args.cpu_track_steps, seed=20260103,
                                        irregular=True, n_dim=2)
    return dat, par, info


def _time_evals(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    return (time.perf_counter() - t0) / steps


def cpu_time_evals(args, steps, warmup):
    """Times nllk + gradient evaluations of a bounded sample of the GPU workload on the host cores.
    Returns the cpu_baseline object (value = all-core figure of the preferred kind)."""
    from oracle import oracle_c, oracle_ref
    from smoothsde_b200 import sharded
    dat, par, info = cpu_problem(args)
    n = info["n"]
    cores = host_cores()
    b1_tracks = max(1, min(args.cpu_b1_tracks, args.cpu_tracks))
    n1 = b1_tracks * args.cpu_track_steps
    sub1 = sharded.shard_rows(dat, 0, n1)[0]
    sample = (f"{args.cpu_tracks} tracks x {args.cpu_track_steps} irregular steps CTCRW d=2 (n={n}), same formulas and "
              f"generator as the GPU workload; {steps} nllk+gradient evaluations after {warmup} warm-up; one-thread "
              f"figure on the first {b1_tracks} tracks (n={n1})")
    out = {"unit": UNIT, "cores": cores, "sample": sample, "rows": n, "steps": steps, "warmup": warmup}
    co = oracle_c.COracle(dat, nthreads=cores)
    s_port = _time_evals(lambda: co.eval(par, True), steps, warmup)
    co1 = oracle_c.COracle(sub1, nthreads=1)
    s_port1 = _time_evals(lambda: co1.eval(par, True), max(1, steps // 4), 1)
    port = {"value": n / s_port, "ms_per_step": s_port * 1e3, "cores": cores,
            "one_thread": {"value": n1 / s_port1, "rows": n1},
            "what": "oracle/oracle_c.c: C restatement of nllk_ctcrw.hpp with a hand-written adjoint, OpenMP over tracks"}
    if oracle_ref.available():
        P = oracle_ref.RefOracleParallel(dat, cores)
        s_ref = _time_evals(lambda: P.eval(par, True), steps, warmup)
        v_ref, g_ref = P.eval(par, True)
        P.close()
        P1 = oracle_ref.RefOracleParallel(sub1, 1)
        s_ref1 = _time_evals(lambda: P1.eval(par, True), max(1, steps // 4), 1)
        P1.close()
        v_port, g_port = co.eval(par, True)
        out.update(value=n / s_ref, ms_per_step=s_ref * 1e3, kind="reference",
                   one_thread={"value": n1 / s_ref1, "rows": n1, "cores": 1},
                   what="the reference's own objective (src/smoothSDE.cpp + src/nllk/nllk_ctcrw.hpp, unmodified) "
                        "compiled against oracle/tmb_shim/TMB.hpp (g++ -O2): eager containers + a reverse-mode AD tape "
                        "re-recorded every evaluation, one track per task on a thread pool. TMB proper (CppAD/TMBad tape "
                        "replay, Eigen) is not installed in this image; per-row arithmetic is the reference's.",
                   port=port, port_agrees=abs(v_port - v_ref) / abs(v_ref))
    else:
        out.update(value=port["value"], ms_per_step=port["ms_per_step"], kind="port", one_thread=port["one_thread"],
                   what=port["what"])
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 20))
    warmup = max(1, min(args.warmup, 3))
    cb = cpu_time_evals(args, steps, warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "CTCRW d=2, tau,nu ~ s(time,k=10), mu fixed 0: bounded sample of BASELINE configs[2] "
                               "(per-row cost does not depend on the number of rows)",
                   "sample": cb["sample"]},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# clocks sampler (NVML)
# ------------------------------------------------------------------------------------------------
class Clocks:
    def __init__(self, index):
        self.samples, self.mem_samples, self.reasons, self.max_mhz, self.max_mem_mhz = [], [], set(), None, None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.max_mem_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_MEM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
                 0x4: "sw_power_cap", 0x80: "hw_power_brake_slowdown"}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                self.mem_samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_MEM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.02)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()
        s, m = sorted(self.samples), sorted(self.mem_samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "mem_mhz": (m[len(m) // 2] if m else None), "mem_max_mhz": self.max_mem_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ------------------------------------------------------------------------------------------------
N1_FILE = os.path.join(ROOT, "profiles", "bench_nllk_n1.json")


def measure(kind, args, ctx):
    """One workload on this rank's GPU: build on the device, warm up, time `steps` evaluations
    device-resident (CUDA events, max over ranks) and end to end through the public host-buffer
    call.  kind: "tracks" (BASELINE configs[2], sharded by track ID) | "single" (configs[3], one
    track sharded along time)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from smoothsde_b200 import devgen, _lib, sharded

    rank, world, local, dev, dist_reduce = ctx["rank"], ctx["world"], ctx["local"], ctx["dev"], ctx["dist_reduce"]
    single = kind == "single"
    ts = tse = None
    comm = sharded.DistComm() if world > 1 else None
    if single:
        n_all = args.tracks * args.track_steps
        segs = 1024                     # the track is simulated as 1024 stitched segments for every world size
        assert segs % world == 0 and n_all % segs == 0
        eng, par, info = devgen.make_ctcrw_device(
            1, n_all // world, seed=20260104, device=local, rank=rank, world=world, sim_tracks=segs // world,
            dist_reduce=dist_reduce, dist_gather=ctx["dist_gather"],
            shard_flags=(_lib.SHARD_NO_PENALTY if rank > 0 else 0), time_shard=world > 1)
        if world > 1:
            ts = sharded.TimeShardedEngine.from_engine_distributed(eng, comm, local)
    else:
        tracks_local = args.tracks // world
        eng, par, info = devgen.make_ctcrw_device(
            tracks_local, args.track_steps, seed=20260103, device=local, rank=rank, world=world,
            dist_reduce=dist_reduce, shard_flags=(_lib.SHARD_NO_PENALTY if rank > 0 else 0))
        if world > 1:
            tse = sharded.TrackShardedEngine.from_engine(eng, comm, local)
    n_local = info["n"]
    n_total = n_local * world
    npar = eng.n_par
    par_dev = torch.as_tensor(par, device=dev)
    out_dev = torch.zeros(npar + 2, dtype=torch.float64, device=dev)
    # a non-default torch stream: the C ABI treats a NULL stream as "the handle's own stream",
    # and torch.cuda.Event only sees work on torch's current stream
    tstream = ts.streams[0] if ts is not None else (tse.stream if tse is not None else torch.cuda.Stream(device=dev))
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0

    def step_device():
        if ts is not None:                 # one track cut along time: 3 stages, 2 all-gathers + 1 all-reduce
            out_dev.copy_(ts.eval_device(par_dev))
            return
        eng.eval_device(par_dev.data_ptr(), out_dev.data_ptr(), 1, stream)
        if world > 1:
            dist.all_reduce(out_dev[:npar + 1])

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    eng.check()                            # no device-side failure
    first = out_dev.cpu().numpy().copy()
    assert np.isfinite(first[:npar + 1]).all(), "non-finite nllk / gradient"

    # ---- device-resident throughput (inputs already in HBM, results stay in HBM) ----
    clocks = Clocks(local)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if ts is not None:
        ts.fallbacks = [0, 0]
    barrier()
    clocks.start()
    ev0.record()
    for _ in range(args.steps):
        step_device()
    ev1.record()
    barrier()
    clk = clocks.stop()
    ms_total = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    dist_reduce(ms_total, "max")
    ms_step = float(ms_total) / args.steps
    value = n_total / (ms_step * 1e-3)
    fallbacks = list(ts.fallbacks) if ts is not None else [0, 0]

    # ---- per-kernel device times (CUDA events in front of every kernel, same stream) ----
    eng.set_profile(True)
    ksum, kcount = {}, 0
    for _ in range(min(args.steps, 10)):
        if ts is not None:
            ts.eval_device(par_dev)        # events between the stages also see the waits for the collectives
        else:
            eng.eval_device(par_dev.data_ptr(), out_dev.data_ptr(), 1, stream)
        torch.cuda.synchronize()
        for nm, ms in eng.last_kernel_times():
            ksum[nm] = ksum.get(nm, 0.0) + ms
        kcount += 1
    eng.set_profile(False)
    kernels = {nm: v / kcount for nm, v in ksum.items()}
    launches_per_step = eng.last_eval_launches

    # ---- end to end through the repo's public call with HOST buffers: Engine.eval (ssde_eval) on one
    #      GPU, TrackShardedEngine.eval / TimeShardedEngine.eval on several: H2D of the parameter
    #      vector, kernels (+ collectives), D2H of [nllk, gradient], host synchronisation, every step
    h2d, d2h = 8 * npar, 8 * (npar + 1)
    api = eng if world == 1 else (ts if ts is not None else tse)
    v = g = None
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        v, g = api.eval(par, order=1)
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    dist_reduce(e2e_s, "max")
    e2e_value = n_total * args.steps / float(e2e_s)
    assert abs(v - first[0]) <= 1e-12 * abs(first[0]), "host-buffer call and device-resident call disagree"

    # ---- the same data set at every N: nllk against the stored 1-GPU value ----
    parity = n1_parity(f"{kind}:{args.tracks}x{args.track_steps}", args, world, rank, first[0], first[1:npar + 1])

    res = dict(kind=kind, eng=eng, info=info, n_local=n_local, n_total=n_total, npar=npar, ms_step=ms_step, value=value,
               clk=clk, kernels=kernels, launches_per_step=launches_per_step, e2e_value=e2e_value, h2d=h2d, d2h=d2h,
               nllk=float(first[0]), parity=parity, fallbacks=fallbacks, launch_info=eng.launch_info())
    return res


def n1_parity(key, args, world, rank, nllk, grad):
    """nllk / gradient of this run against the stored 1-GPU result of the same data set
    (profiles/bench_nllk_n1.json; long gradients are stored as their first 64 entries + l2 norm)."""
    import numpy as np
    stored = json.load(open(N1_FILE)) if os.path.exists(N1_FILE) else {}
    grad = np.asarray(grad, dtype=float)
    out = None
    if key in stored:
        st = stored[key]
        rel = abs(float(nllk) - st["nllk"]) / abs(st["nllk"])
        gs = np.asarray(st["grad"])
        head = grad[:gs.size]
        gerr = float(np.max(np.abs(head - gs) / np.maximum(np.abs(gs), 1e-3 * np.abs(gs).max())))
        nerr = abs(float(np.linalg.norm(grad)) - st.get("grad_l2", float(np.linalg.norm(gs)))) / st.get("grad_l2", float(np.linalg.norm(gs)))
        ok = bool(rel <= 1e-10 and gerr <= 1e-7 and nerr <= 1e-7)
        if not ok and rank == 0:
            print(f"PARITY FAILURE {key}: {world}-GPU result differs from the stored 1-GPU result: "
                  f"{rel:.3e} / {gerr:.3e} / {nerr:.3e}", file=sys.stderr, flush=True)
        out = {"nllk_rel": rel, "grad_rel": gerr, "grad_l2_rel": nerr, "ok": ok, "tolerance": "1e-10 nllk, 1e-7 gradient"}
    if args.write_n1 and world == 1 and rank == 0:
        stored[key] = {"nllk": float(nllk), "grad": [float(x) for x in grad[:64]], "grad_l2": float(np.linalg.norm(grad))}
        os.makedirs(os.path.dirname(N1_FILE), exist_ok=True)
        json.dump(stored, open(N1_FILE, "w"), indent=1)
    return out


def measure_ou(name, n_tracks, n_steps, args, ctx):
    """BM / OU fused path: OU with a random intercept per track (BASELINE configs[1] and the OU half of
    configs[4]), tracks sharded over the ranks, one all-reduce of [nllk, gradient] per evaluation."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from smoothsde_b200 import devgen, _lib, sharded
    rank, world, local, dev, dist_reduce = ctx["rank"], ctx["world"], ctx["local"], ctx["dev"], ctx["dist_reduce"]
    if n_tracks % world != 0:
        return {"skipped": f"{n_tracks} tracks do not divide over {world} ranks"}
    eng, par, info = devgen.make_ou_device(n_tracks // world, n_steps, seed=20260102, device=local, rank=rank, world=world,
                                           shard_flags=(_lib.SHARD_NO_PENALTY if rank > 0 else 0))
    n_total = info["n"] * world
    npar = eng.n_par
    tse = sharded.TrackShardedEngine.from_engine(eng, sharded.DistComm() if world > 1 else sharded.SoloComm(), local)
    torch.cuda.set_stream(tse.stream)
    stream = tse.stream.cuda_stream
    par_dev = torch.as_tensor(par, device=dev)
    out_dev = torch.zeros(npar + 2, dtype=torch.float64, device=dev)

    def step_device():
        eng.eval_device(par_dev.data_ptr(), out_dev.data_ptr(), 1, stream)
        if world > 1:
            dist.all_reduce(out_dev[:npar + 1])

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    eng.check()
    first = out_dev.cpu().numpy().copy()
    assert np.isfinite(first[:npar + 1]).all()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step_device()
    ev1.record()
    barrier()
    ms_total = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    dist_reduce(ms_total, "max")
    ms_step = float(ms_total) / args.steps
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        v, g = tse.eval(par, order=1)
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    dist_reduce(e2e_s, "max")
    eng.set_profile(True)
    eng.eval_device(par_dev.data_ptr(), out_dev.data_ptr(), 1, stream)
    torch.cuda.synchronize()
    kernels = dict(eng.last_kernel_times())
    eng.set_profile(False)
    launches = eng.last_eval_launches
    # the one-pass data-term Hessian X'WX + lambda S of this rank's rows (the Laplace inner problem's H_bb)
    p_theta = info["p_fe"] + info["p_re"]
    hess_ms = None
    try:
        Hbuf = torch.zeros((p_theta, p_theta), dtype=torch.float64, device=dev)
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        eng.hess_theta_device(par_dev.data_ptr(), Hbuf.data_ptr(), stream)
        torch.cuda.synchronize()
        h0.record()
        for _ in range(3):
            eng.hess_theta_device(par_dev.data_ptr(), Hbuf.data_ptr(), stream)
        h1.record()
        torch.cuda.synchronize()
        hess_ms = h0.elapsed_time(h1) / 3
        launches_h = eng.last_eval_launches
        del Hbuf
    except Exception as e:                       # noqa: BLE001 -- reported, not fatal for the throughput line
        hess_ms = f"failed: {e}"
        launches_h = 0
    parity = n1_parity(f"{name}:{n_tracks}x{n_steps}", args, world, rank, first[0], first[1:npar + 1])
    stored_b = info["stored_bytes_per_obs"]       # design values (mu / tau blocks stored once) + dt + obs + flag
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = float(json.load(open(peaks_path))["hbm_gbs"]) if os.path.exists(peaks_path) else 6650.0
    kmain = max(kernels, key=kernels.get)
    res = {"workload": f"OU d=1, {n_tracks} tracks x {n_steps} steps (n={n_total}), mu,tau ~ s(time,k=10) + s(ID,bs='re'), kappa ~ 1; "
                       f"p_re = {info['p_re']}", "ms_per_step": ms_step, "value": n_total / (ms_step * 1e-3), "unit": UNIT,
           "e2e": {"value": n_total * args.steps / float(e2e_s), "unit": UNIT, "h2d_bytes_per_step": 8 * npar,
                   "d2h_bytes_per_step": 8 * (npar + 1), "call": "TrackShardedEngine.eval"},
           "kernels_ms": kernels, "gpu_launches": launches * args.steps + 4 * launches_h, "nllk": float(first[0]), "parity_vs_n1": parity,
           "hessian_onepass_ms": hess_ms,
           "roofline": {"kernel": kmain, "launch_ms": kernels[kmain], "stored_bytes_per_obs": stored_b,
                        "alg_bytes_per_obs": devgen.alg_bytes_per_obs(1, 3, 23),
                        "dram_frac": stored_b * info["n"] / (kernels[kmain] * 1e-3) / 1e9 / peak,
                        "frac": devgen.alg_bytes_per_obs(1, 3, 23) * info["n"] / (kernels[kmain] * 1e-3) / 1e9 / peak,
                        "note": "dram_frac: bytes the layout stores (design read ONCE) / launch time / measured peak; frac: SURVEY 8(d) "
                                "CSR accounting (design read twice), > 1 because the kernel needs one pass"}}
    eng.close()
    torch.cuda.empty_cache()
    return res


def release(res):
    import torch
    res["eng"].close()
    res.pop("eng")
    res.pop("info")
    torch.cuda.empty_cache()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return
    import torch
    import torch.distributed as dist
    from smoothsde_b200 import devgen

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert args.tracks % world == 0, "tracks must divide by the number of ranks"

    def dist_reduce(t, op):
        if world > 1:
            dist.all_reduce(t, op={"min": dist.ReduceOp.MIN, "max": dist.ReduceOp.MAX, "sum": dist.ReduceOp.SUM}[op])
        return t

    def dist_gather(t):
        out = torch.empty((world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t.contiguous())
        return out

    ctx = dict(rank=rank, world=world, local=local, dev=dev, dist_reduce=dist_reduce, dist_gather=dist_gather)
    head_kind = "single" if args.workload == "single" else "tracks"
    r = measure(head_kind, args, ctx)
    info, n_local, npar, kernels = r["info"], r["n_local"], r["npar"], r["kernels"]

    # ---- roofline of the dominant kernel ----
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = float(json.load(open(peaks_path))["hbm_gbs"]); peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak = 6650.0; peak_src = "fallback (B200_PROFILING.md)"
    # B_alg = (8d + 8 + 4) + 2 [12 nnz/n + 8 n_par] = 620 B/obs per evaluation.  Per
    # launch: the forward kernel owns the data term and one design pass
    # (28 + 296 = 324 B/obs), the adjoint kernel the second design pass (296 B/obs).
    nnz_row = info["nnz"] / n_local
    b_alg = devgen.alg_bytes_per_obs(info["n_dim"], info["n_par"], nnz_row)
    b_pass = 12 * nnz_row + 8 * info["n_par"]
    b_kernel = {"ctcrw_fwd": (8 * info["n_dim"] + 8 + 4) + b_pass, "ctcrw_bwd": b_pass}
    dev_ms = sum(kernels.values())
    dom = max((k for k in kernels if k in b_kernel), key=kernels.get)
    achieved = b_kernel[dom] * n_local / (kernels[dom] * 1e-3) / 1e9
    traffic = dram_frac = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):                     # dram bytes of one launch from the committed ncu capture
        tj = json.load(open(tpath))
        k = tj["kernels"].get(dom)
        if k:
            traffic = (k["dram_bytes_read"] + k["dram_bytes_write"]) / tj["rows"] * n_local
            dram_frac = traffic / (kernels[dom] * 1e-3) / 1e9 / peak
    ach_eval = b_alg * n_local / (dev_ms * 1e-3) / 1e9
    by_kernel = {}                                # both scan kernels side by side (the dominant one is the headline)
    try:
        tk = json.load(open(tpath))["kernels"] if os.path.exists(tpath) else {}
        trows = json.load(open(tpath))["rows"] if os.path.exists(tpath) else 1
        for kn, bk in b_kernel.items():
            if kn not in kernels:
                continue
            sec = kernels[kn] * 1e-3
            ent = {"launch_ms": kernels[kn], "frac": bk * n_local / sec / 1e9 / peak}
            if kn in tk:
                ent["dram_frac"] = (tk[kn]["dram_bytes_read"] + tk[kn]["dram_bytes_write"]) / trows * n_local / sec / 1e9 / peak
            by_kernel[kn] = ent
    except Exception:                             # noqa: BLE001 -- an extra, never worth losing the line for
        by_kernel = {}
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": traffic, "dram_frac": dram_frac, "peak_source": peak_src, "kernel": dom,
        "note": "frac = ALGORITHMIC bytes (SURVEY 8(d) CSR accounting) / launch time / peak; dram_frac = DRAM bytes "
                "actually moved (ncu capture of this kernel, scaled by rows) / launch time / peak -- the packed layout "
                "stores fewer bytes than the CSR accounting, so dram_frac is the memory system's real utilisation",
        "alg_bytes_per_launch": b_kernel[dom] * n_local, "launch_ms": kernels[dom],
        "traffic_source": "profiles/ncu_traffic.json (ncu --set full capture of this workload, scaled by rows per GPU)",
        "kernels_ms": kernels, "dominant_share": kernels[dom] / dev_ms,
        "alg_bytes_per_obs_by_kernel": b_kernel, "by_kernel": by_kernel,
        "evaluation": {"alg_bytes_per_obs": b_alg, "achieved": ach_eval, "frac": ach_eval / peak,
                       "note": "whole evaluation: 620 B/obs (SURVEY 8(d)) x rows on this GPU / summed kernel time"},
    }
    single = head_kind == "single"
    n_total = r["n_total"]
    line = {
        "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": r["ms_step"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": (f"CTCRW d=2, ONE track of {n_total} irregular steps, tau,nu ~ s(time,k=10), mu fixed 0 "
                                f"(BASELINE configs[3])" if single else
                                f"CTCRW d=2, {args.tracks} tracks x {args.track_steps} irregular steps "
                                f"(n={n_total}), tau,nu ~ s(time,k=10), mu fixed 0 (BASELINE configs[2])"),
                   "sharding": (f"track cut along time over {world} rank(s): 2 NCCL all-gathers of one scan element + 1 all-reduce "
                                f"of {npar + 2} doubles per evaluation" if single else
                                f"tracks split over {world} rank(s), one NCCL all-reduce of {npar + 1} doubles per evaluation"),
                   "same_data_at_every_n": "random draws are seeded per group of 64 tracks (devgen.GroupedRNG), so 1/2/4/8 "
                                           "GPUs evaluate the identical data set; parity_vs_n1 checks nllk/gradient against "
                                           "the stored 1-GPU result (profiles/bench_nllk_n1.json)",
                   "l2": "inputs per GPU (>= 3 GB) are far larger than the 126 MB L2; no flush needed"},
        "clocks": r["clk"],
        "e2e": {"value": r["e2e_value"], "unit": UNIT, "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": r["d2h"],
                "call": "Engine.eval (ssde_eval)" if world == 1 else ("TimeShardedEngine.eval" if single else "TrackShardedEngine.eval")},
        "gpu_launches": r["launches_per_step"] * args.steps,
        "roofline": roofline,
        "nllk": r["nllk"],
        "parity_vs_n1": r["parity"],
        "launch_info": r["launch_info"],
    }
    release(r)
    if args.workload == "both":
        # BASELINE configs[3]: ONE track of the same total length, sharded along time
        r2 = measure("single", args, ctx)
        k2 = r2["kernels"]
        line["single"] = {
            "workload": f"CTCRW d=2, ONE track of {r2['n_total']} irregular steps (BASELINE configs[3]), time-sharded over {world} rank(s)",
            "ms_per_step": r2["ms_step"], "value": r2["value"], "unit": UNIT,
            "e2e": {"value": r2["e2e_value"], "unit": UNIT, "h2d_bytes_per_step": r2["h2d"], "d2h_bytes_per_step": r2["d2h"],
                    "call": "Engine.eval (ssde_eval)" if world == 1 else "TimeShardedEngine.eval"},
            "collectives_per_eval": ({"all_gather": 2, "all_reduce": 1} if world > 1 else {}),
            "full_shard_fallbacks": {"forward_stage3": r2["fallbacks"][0], "adjoint_stage4": r2["fallbacks"][1],
                                     "of_evaluations": args.steps},
            "kernels_ms": k2, "gpu_launches": r2["launches_per_step"] * args.steps, "nllk": r2["nllk"],
            "parity_vs_n1": r2["parity"], "clocks": r2["clk"],
        }
        line["gpu_launches"] += r2["launches_per_step"] * args.steps
        release(r2)
    if args.workload == "both" and not args.no_extra_configs:
        # the other BASELINE configs on the BM / OU path (driver-visible at every N)
        line["configs"] = {
            "configs[4] OU half": measure_ou("ou", 4096, 25000, args, ctx),
            "configs[1]": measure_ou("ou", 64, 100000, args, ctx),
        }
        for c in line["configs"].values():
            line["gpu_launches"] += c.get("gpu_launches", 0)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_time_evals(args, 3, 1)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
