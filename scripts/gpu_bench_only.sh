mkdir -p gpurun_out
TAG=${1:-x}
timeout 600 python -m pytest tests -m gpu -x -q -k "golden or edge" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_full.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_bench_full.log
tail -n 2 gpurun_out/${TAG}_pytest.log
