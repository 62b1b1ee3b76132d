mkdir -p gpurun_out
T=r26
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_full.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_bench_full.log
tail -n 30 gpurun_out/${T}_pytest.log; tail -n 2 gpurun_out/${T}_bench_full.log | cut -c1-2500
