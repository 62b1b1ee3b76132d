mkdir -p gpurun_out
T=r17
nvidia-smi -L > gpurun_out/${T}_smi.log
timeout 1200 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/${T}_bench2.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_bench2.log
tail -n 30 gpurun_out/${T}_pytest.log; tail -n 3 gpurun_out/${T}_bench2.log | cut -c1-1500
