"""Why are the per-GPU kernels slower under torchrun than in a single process?  Times the same per-rank
problem (512 tracks x 1e5) (a) before torch.distributed is initialised, (b) after NCCL init, (c) after the
first collective, (d) built with the cross-rank reductions of bench.py."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from smoothsde_b200 import devgen

rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
T = 1024 // max(world, 1)


def timeit(eng, par, tag):
    par_dev = torch.as_tensor(par, device=dev)
    out = torch.zeros(eng.n_par + 2, dtype=torch.float64, device=dev)
    st = torch.cuda.Stream(device=dev)
    eng.set_profile(True)
    acc = {}
    for i in range(8):
        eng.eval_device(par_dev.data_ptr(), out.data_ptr(), 1, st.cuda_stream)
        torch.cuda.synchronize()
        if i >= 3:
            for nm, ms in eng.last_kernel_times():
                acc[nm] = acc.get(nm, 0.0) + ms / 5
    eng.set_profile(False)
    print(f"[rank {rank}] {tag}: fwd {acc['ctcrw_fwd']:.3f} bwd {acc['ctcrw_bwd']:.3f}", flush=True)


if not os.environ.get("PROBE_BENCH"):
    eng, par, info = devgen.make_ctcrw_device(T, 100000, seed=20260103, device=local, rank=rank, world=world)
    timeit(eng, par, "(a) local build, no torch.distributed")
if world > 1 and not os.environ.get("PROBE_BENCH"):
    dist.init_process_group("nccl", device_id=dev)
    timeit(eng, par, "(b) after NCCL init")
    t = torch.ones(4, device=dev, dtype=torch.float64)
    dist.all_reduce(t)
    torch.cuda.synchronize()
    timeit(eng, par, "(c) after the first all_reduce")
    eng.close()
    del eng, info
    torch.cuda.empty_cache()

    def dist_reduce(x, op):
        dist.all_reduce(x, op={"min": dist.ReduceOp.MIN, "max": dist.ReduceOp.MAX, "sum": dist.ReduceOp.SUM}[op])
        return x
    eng, par, info = devgen.make_ctcrw_device(T, 100000, seed=20260103, device=local, rank=rank, world=world, dist_reduce=dist_reduce)
    timeit(eng, par, "(d) built with cross-rank reductions")
    dist.barrier()
    dist.destroy_process_group()

if world > 1 and os.environ.get("PROBE_BENCH"):
    # replay bench.py's sequence stage by stage
    import bench as B
    from smoothsde_b200 import sharded, _lib
    dist.init_process_group("nccl", device_id=dev)

    def dist_reduce(x, op):
        dist.all_reduce(x, op={"min": dist.ReduceOp.MIN, "max": dist.ReduceOp.MAX, "sum": dist.ReduceOp.SUM}[op])
        return x
    eng, par, info = devgen.make_ctcrw_device(T, 100000, seed=20260103, device=local, rank=rank, world=world, dist_reduce=dist_reduce,
                                              shard_flags=(_lib.SHARD_NO_PENALTY if rank > 0 else 0))
    timeit(eng, par, "(e) bench build incl. shard flags")
    tse = sharded.TrackShardedEngine.from_engine(eng, sharded.DistComm(), local)
    timeit(eng, par, "(f) after TrackShardedEngine.from_engine")
    torch.cuda.set_stream(tse.stream)
    stream = tse.stream.cuda_stream
    npar = eng.n_par
    par_dev = torch.as_tensor(par, device=dev)
    out_dev = torch.zeros(npar + 2, dtype=torch.float64, device=dev)
    for _ in range(3):
        eng.eval_device(par_dev.data_ptr(), out_dev.data_ptr(), 1, stream)
        dist.all_reduce(out_dev[:npar + 1])
    torch.cuda.synchronize(); dist.barrier()
    timeit(eng, par, "(g) after warm-up steps with all_reduce on the engine's stream")
    clocks = B.Clocks(local)
    clocks.start()
    for _ in range(20):
        eng.eval_device(par_dev.data_ptr(), out_dev.data_ptr(), 1, stream)
        dist.all_reduce(out_dev[:npar + 1])
    torch.cuda.synchronize(); dist.barrier()
    print(clocks.stop())
    timeit(eng, par, "(h) after the timed loop with the NVML sampler")
    eng.set_profile(True)
    acc = {}
    for i in range(10):
        eng.eval_device(par_dev.data_ptr(), out_dev.data_ptr(), 1, stream)
        torch.cuda.synchronize()
        for nm, ms in eng.last_kernel_times():
            acc[nm] = acc.get(nm, 0.0) + ms / 10
    eng.set_profile(False)
    print(f"[rank {rank}] (i) bench-style profile loop on tse.stream: fwd {acc['ctcrw_fwd']:.3f} bwd {acc['ctcrw_bwd']:.3f}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
