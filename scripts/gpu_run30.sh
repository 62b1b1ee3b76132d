mkdir -p gpurun_out
T=r30
timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload single --steps 10 --warmup 3 > gpurun_out/${T}_single2.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_single2.log
tail -5 gpurun_out/${T}_pytest.log; tail -n 2 gpurun_out/${T}_single2.log | cut -c1-2600
