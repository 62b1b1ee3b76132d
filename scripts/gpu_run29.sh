mkdir -p gpurun_out
T=r29
timeout 600 python bench.py --workload single --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_single1.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_single1.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload single --steps 10 --warmup 3 > gpurun_out/${T}_single2.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_single2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/${T}_tracks2.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_tracks2.log
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest.log
for f in single1 single2 tracks2; do tail -n 2 gpurun_out/${T}_$f.log | cut -c1-400; done; tail -3 gpurun_out/${T}_pytest.log
