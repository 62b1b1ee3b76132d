mkdir -p gpurun_out
T=r31
timeout 400 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest.log
tail -5 gpurun_out/${T}_pytest.log
