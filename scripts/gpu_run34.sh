mkdir -p gpurun_out
T=r34
timeout 900 python -m pytest tests/test_gpu_dense.py -m gpu -q > gpurun_out/${T}_dense.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_dense.log
timeout 1200 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_dense.py > gpurun_out/${T}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/${T}_bench.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_bench.log
tail -40 gpurun_out/${T}_dense.log; tail -5 gpurun_out/${T}_pytest.log; tail -n 2 gpurun_out/${T}_bench.log | cut -c1-400
