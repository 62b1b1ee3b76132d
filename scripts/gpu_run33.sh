mkdir -p gpurun_out
T=r33
K='regex:ctcrw_fwd|ctcrw_bwd'
timeout 900 ncu --set full --clock-control none --import-source on -k "$K" --launch-skip 6 --launch-count 2 -f -o gpurun_out/${T}_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_full.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_ncu_full.log
tail -n 3 gpurun_out/${T}_ncu_full.log | cut -c1-300
ls -la gpurun_out/
