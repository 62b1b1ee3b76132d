mkdir -p gpurun_out
T=r25
SSDE_LIB_SUFFIX=_stats python scripts/stats_run.py 1024 100000 > gpurun_out/${T}_stats.log 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_full.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_bench_full.log
tail -8 gpurun_out/${T}_stats.log; tail -n 5 gpurun_out/${T}_pytest.log; tail -n 2 gpurun_out/${T}_bench_full.log | cut -c1-2500
