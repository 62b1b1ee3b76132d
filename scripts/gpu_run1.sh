mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/r1_smi.log 2>&1
nproc >> gpurun_out/r1_smi.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r1_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r1_smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 --tracks 64 --track-steps 100000 > gpurun_out/r1_bench_small.log 2>&1; echo "rc=$?" >> gpurun_out/r1_bench_small.log
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r1_bench_full.log 2>&1; echo "rc=$?" >> gpurun_out/r1_bench_full.log
tail -3 gpurun_out/r1_*.log
