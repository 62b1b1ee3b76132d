mkdir -p gpurun_out
T=r35
timeout 900 python -m pytest tests/test_decay.py tests/test_gpu_dense.py -m gpu -q > gpurun_out/${T}_new.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_new.log
timeout 1200 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_dense.py --deselect tests/test_decay.py > gpurun_out/${T}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/${T}_bench.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_bench.log
grep -v Warning gpurun_out/${T}_new.log | tail -30; tail -3 gpurun_out/${T}_pytest.log; tail -n 2 gpurun_out/${T}_bench.log | cut -c1-400
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r35_bench.log') if x.startswith('{')][-1]
d=json.loads(l); print(d['ms_per_step'], d['roofline']['kernels_ms'])
PY
