# round-2 call A: full GPU suite on the aliased layout + kernel variants side by side
mkdir -p gpurun_out
T=${1:-ra}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest.log
grep -v Warn gpurun_out/${T}_pytest.log | tail -5
SSDE_NO_ALIAS=1 bash scripts/gpu_tune.sh ${T}_noalias v0
bash scripts/gpu_tune.sh ${T} v0 v1 v2 v3
