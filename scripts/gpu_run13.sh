mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r13_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r13_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r13_bench_full.log 2>&1; echo "rc=$?" >> gpurun_out/r13_bench_full.log
K='regex:ctcrw_fwd|ctcrw_bwd'
timeout 900 ncu --set full --clock-control none --import-source on -k "$K" --launch-skip 6 --launch-count 2 -f -o gpurun_out/r13_full python bench.py --steps 2 --warmup 3 --tracks 64 --track-steps 100000 --no-cpu-baseline > gpurun_out/r13_ncu_full.log 2>&1; echo "rc=$?" >> gpurun_out/r13_ncu_full.log
tail -n 4 gpurun_out/r13_pytest.log
