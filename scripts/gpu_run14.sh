mkdir -p gpurun_out
T=r14
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/${T}_smi.log 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench_full.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_bench_full.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${T}_bench_ref.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_bench_ref.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_launch.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_ncu_launch.log
K='regex:ctcrw_fwd|ctcrw_bwd'
timeout 900 ncu --set full --clock-control none --import-source on -k "$K" --launch-skip 6 --launch-count 2 -f -o gpurun_out/${T}_full python bench.py --steps 2 --warmup 3 --tracks 64 --track-steps 100000 --no-cpu-baseline > gpurun_out/${T}_ncu_full.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_ncu_full.log
tail -n 4 gpurun_out/${T}_pytest.log; tail -n 3 gpurun_out/${T}_bench_full.log
