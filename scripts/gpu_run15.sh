mkdir -p gpurun_out
T=r15
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_full.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_bench_full.log
K='regex:ssde|ctcrw|finalize|gather_theta|reduce_tiles|sde_fused'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_launch.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_ncu_launch.log
K='regex:ctcrw_fwd|ctcrw_bwd'
timeout 900 ncu --set full --clock-control none --import-source on -k "$K" --launch-skip 6 --launch-count 2 -f -o gpurun_out/${T}_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_full.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_ncu_full.log
tail -n 25 gpurun_out/${T}_pytest.log; tail -n 3 gpurun_out/${T}_bench_full.log | cut -c1-600
