# round-2 call B: full GPU suite (likelihood summed in the adjoint kernel, aliased BM/OU layout),
# variants side by side, ncu capture of the CTCRW kernels with per-instruction counters
mkdir -p gpurun_out
T=${1:-rb}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest.log
grep -v Warn gpurun_out/${T}_pytest.log | tail -5
bash scripts/gpu_tune.sh ${T} v1 v4
SSDE_LIB_SUFFIX=_v4 NCU_SKIP=6 timeout 600 bash scripts/ncu_capture.sh ${T}_ctcrw "ctcrw_fwd|ctcrw_bwd" 2 python bench.py --workload tracks --steps 2 --warmup 3 --no-cpu-baseline
head -60 gpurun_out/${T}_ctcrw.md | grep -i "time\|issue\|dram\|stall\|regs"
timeout 300 python scripts/ou_run.py 4096 25000 > gpurun_out/${T}_ou.log 2>&1; tail -3 gpurun_out/${T}_ou.log; SSDE_NO_ALIAS=1 timeout 300 python scripts/ou_run.py 4096 25000 > gpurun_out/${T}_ou_noalias.log 2>&1; tail -1 gpurun_out/${T}_ou_noalias.log
