mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3_pytest.log
timeout 600 compute-sanitizer --tool memcheck python tests/gpu_smoke_small.py > gpurun_out/r3_memcheck.log 2>&1; echo "rc=$?" >> gpurun_out/r3_memcheck.log
timeout 600 python bench.py --steps 5 --warmup 3 --tracks 64 --track-steps 100000 --no-cpu-baseline > gpurun_out/r3_bench_small.log 2>&1; echo "rc=$?" >> gpurun_out/r3_bench_small.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r3_bench_full.log 2>&1; echo "rc=$?" >> gpurun_out/r3_bench_full.log
tail -n 5 gpurun_out/r3_pytest.log gpurun_out/r3_memcheck.log
