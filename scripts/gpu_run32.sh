mkdir -p gpurun_out
T=r32
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_smoke.log
timeout 600 python bench.py > gpurun_out/${T}_bench.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_bench.log
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${T}_bench_ref.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_bench.log 2>&1
tail -5 gpurun_out/${T}_pytest.log; tail -3 gpurun_out/${T}_smoke.log; tail -n 1 gpurun_out/${T}_bench.log | cut -c1-3000
