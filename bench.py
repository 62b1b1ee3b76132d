#!/usr/bin/env python
"""Benchmark of the hot path: joint nllk + gradient evaluations of a 1e8-observation CTCRW model.

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference          # CPU arm (oracle port of the TMB objective)

A "step" is one evaluation of the joint penalised nllk and its gradient (what obj$fn + obj$gr do
in SDE$fit(), R/sde.R:694-697) over the whole synthetic data set, which is resident in HBM.
Workload (BASELINE.json configs[2], the CTCRW 1e8-observation case the metric is quoted on):
1024 simulated tracks x 1e5 irregular steps, d = 2, tau, nu ~ s(time, k = 10), mu fixed at 0.
With N > 1 GPUs the tracks are split across ranks (strong scaling: the total stays 1.024e8
rows) and the packed [nllk, gradient] vector is summed with one NCCL all-reduce per evaluation.
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
    ap.add_argument("--cpu-tracks", type=int, default=64)
    ap.add_argument("--cpu-track-steps", type=int, default=10000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="tracks", choices=["tracks", "single"],
                    help="tracks: BASELINE configs[2] (default, the headline); single: configs[3], ONE track of "
                         "tracks*track_steps rows, sharded along time over the ranks")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port (dense 2d x 2d Kalman recursion + hand adjoint, OpenMP over tracks)
# ------------------------------------------------------------------------------------------------
def cpu_problem(args):
    from smoothsde_b200 import synth
    dat, par, info = synth.make_problem("CTCRW", args.cpu_tracks, args.cpu_track_steps, seed=20260103,
                                        irregular=True, n_dim=2)
    return dat, par, info


def cpu_time_evals(args, steps, warmup):
    """Returns (obs*eval/s, cores, sample description)."""
    from oracle import oracle_c
    dat, par, info = cpu_problem(args)
    cores = min(oracle_c.max_threads(), os.cpu_count() or 1)
    co = oracle_c.COracle(dat, nthreads=cores)
    for _ in range(max(warmup, 1)):
        co.eval(par, True)
    t0 = time.perf_counter()
    for _ in range(steps):
        co.eval(par, True)
    dt = time.perf_counter() - t0
    sample = (f"{args.cpu_tracks} tracks x {args.cpu_track_steps} steps CTCRW d=2 (n={info['n']}), "
              f"same formulas as the GPU workload, {steps} nllk+gradient evaluations")
    return info["n"] * steps / dt, cores, sample, dt / steps * 1e3


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 20))
    val, cores, sample, ms = cpu_time_evals(args, steps, min(args.warmup, 2))
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": min(args.warmup, 2), "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "CTCRW d=2, tau,nu ~ s(time,k=10), bounded sample of BASELINE configs[2]",
                   "sample": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "note": "C restatement of src/nllk/nllk_ctcrw.hpp with a hand-written adjoint; "
                                 "TMB itself cannot be built in this image (no R/TMB/Eigen)"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# clocks sampler (NVML)
# ------------------------------------------------------------------------------------------------
class Clocks:
    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
                 0x4: "sw_power_cap", 0x80: "hw_power_brake_slowdown"}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
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
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ------------------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return
    import numpy as np
    import torch
    import torch.distributed as dist
    from smoothsde_b200 import devgen, _lib

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

    single = args.workload == "single"
    ts = None
    if single:
        from smoothsde_b200 import sharded
        n_all = args.tracks * args.track_steps
        assert n_all % (world * 128) == 0
        eng, par, info = devgen.make_ctcrw_device(
            1, n_all // world, seed=20260104, device=local, rank=rank, world=world, sim_tracks=128,
            dist_reduce=dist_reduce, shard_flags=(_lib.SHARD_NO_PENALTY if rank > 0 else 0), time_shard=world > 1)
        if world > 1:
            ts = sharded.TimeShardedEngine.from_engine_distributed(eng, sharded.DistComm(), local)
    else:
        tracks_local = args.tracks // world
        eng, par, info = devgen.make_ctcrw_device(
            tracks_local, args.track_steps, seed=20260103, device=local, rank=rank, world=world,
            dist_reduce=dist_reduce, shard_flags=(_lib.SHARD_NO_PENALTY if rank > 0 else 0))
    n_local = info["n"]
    n_total = n_local * world
    npar = eng.n_par
    par_dev = torch.as_tensor(par, device=dev)
    out_dev = torch.zeros(npar + 2, dtype=torch.float64, device=dev)
    par_host = torch.as_tensor(par).pin_memory()
    out_host = torch.zeros(npar + 2, dtype=torch.float64).pin_memory()
    # a non-default torch stream: the C ABI treats a NULL stream as "the handle's own stream",
    # and torch.cuda.Event only sees work on torch's current stream
    tstream = ts.streams[0] if ts is not None else torch.cuda.Stream(device=dev)
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
    # sanity: finite objective, no device-side failure
    eng.check()
    first = out_dev.cpu().numpy().copy()
    assert np.isfinite(first[:npar + 1]).all(), "non-finite nllk / gradient"

    # ---- device-resident throughput (inputs already in HBM, results stay in HBM) ----
    clocks = Clocks(local)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
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

    # ---- end to end through the public call with HOST buffers ----
    h2d, d2h = 8 * npar, 8 * (npar + 1)
    barrier()
    t0 = time.perf_counter()
    if world == 1:
        for _ in range(args.steps):
            v, g = eng.eval(par, order=1)           # ssde_eval: H2D par, kernels, D2H nllk+grad, sync
    elif ts is not None:
        for _ in range(args.steps):
            v, g = ts.eval(par)                     # H2D par, 3 stages + collectives, D2H nllk+grad, sync
    else:
        for _ in range(args.steps):
            par_dev.copy_(par_host, non_blocking=True)
            step_device()
            out_host.copy_(out_dev, non_blocking=True)
            torch.cuda.synchronize()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    dist_reduce(e2e_s, "max")
    e2e_value = n_total * args.steps / float(e2e_s)

    # ---- roofline ----
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
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):                     # dram bytes of one launch from the committed ncu capture
        tj = json.load(open(tpath))
        k = tj["kernels"].get(dom)
        if k:
            traffic = (k["dram_bytes_read"] + k["dram_bytes_write"]) / tj["rows"] * n_local
    ach_eval = b_alg * n_local / (dev_ms * 1e-3) / 1e9
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": traffic, "peak_source": peak_src, "kernel": dom,
        "alg_bytes_per_launch": b_kernel[dom] * n_local, "launch_ms": kernels[dom],
        "traffic_source": "profiles/ncu_traffic.json (ncu --set full capture of this workload, scaled by rows per GPU)",
        "kernels_ms": kernels, "dominant_share": kernels[dom] / dev_ms,
        "alg_bytes_per_obs_by_kernel": b_kernel,
        "evaluation": {"alg_bytes_per_obs": b_alg, "achieved": ach_eval, "frac": ach_eval / peak,
                       "note": "whole evaluation: 620 B/obs (SURVEY 8(d)) x rows on this GPU / summed kernel time"},
    }

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": (f"CTCRW d=2, ONE track of {n_total} irregular steps, tau,nu ~ s(time,k=10), mu fixed 0 "
                                f"(BASELINE configs[3])" if single else
                                f"CTCRW d=2, {args.tracks} tracks x {args.track_steps} irregular steps "
                                f"(n={n_total}), tau,nu ~ s(time,k=10), mu fixed 0 (BASELINE configs[2])"),
                   "sharding": (f"track cut along time over {world} rank(s): 2 NCCL all-gathers of one scan element + 1 all-reduce "
                                f"of {npar + 2} doubles per evaluation" if single else
                                f"tracks split over {world} rank(s), one NCCL all-reduce of {npar + 1} doubles per evaluation"),
                   "l2": "inputs per GPU (>= 3 GB) are far larger than the 126 MB L2; no flush needed"},
        "clocks": clk,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": launches_per_step * args.steps,
        "roofline": roofline,
        "nllk": float(first[0]),
        "launch_info": eng.launch_info(),
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        val, cores, sample, ms = cpu_time_evals(args, 5, 1)
        line["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
