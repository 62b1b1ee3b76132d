mkdir -p gpurun_out
T=r28
timeout 2400 python scripts/bench_configs.py > gpurun_out/${T}_configs.jsonl 2> gpurun_out/${T}_configs.err; echo "rc=$?" >> gpurun_out/${T}_configs.err
tail -3 gpurun_out/${T}_configs.err; cut -c1-420 gpurun_out/${T}_configs.jsonl
