mkdir -p gpurun_out
T=r18
timeout 1500 python -m pytest tests/test_gpu_laplace.py -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -n 40 gpurun_out/${T}_pytest.log
