mkdir -p gpurun_out
T=${1:-val}
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_smoke.log
timeout 600 python bench.py > gpurun_out/${T}_bench.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_bench.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_bench.log 2>&1
K='regex:ctcrw_fwd|ctcrw_bwd'
timeout 900 ncu --set full --clock-control none --import-source on -k "$K" --launch-skip 6 --launch-count 2 -f -o gpurun_out/${T}_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_full.log 2>&1
timeout 900 python scripts/bench_configs.py > gpurun_out/${T}_configs.jsonl 2> gpurun_out/${T}_configs.err
grep -v Warn gpurun_out/${T}_pytest.log | tail -4; tail -3 gpurun_out/${T}_smoke.log; tail -n 2 gpurun_out/${T}_bench.log | cut -c1-300
