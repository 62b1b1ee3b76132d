mkdir -p gpurun_out
T=r27
timeout 1800 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_laplace.py -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -n 30 gpurun_out/${T}_pytest.log
