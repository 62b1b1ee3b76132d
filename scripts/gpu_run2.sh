mkdir -p gpurun_out
K='regex:ctcrw|linpred|finalize|gather_theta|sde_fused'
timeout 900 ncu --set full --clock-control none --import-source on -k "$K" --launch-skip 18 --launch-count 6 -f -o gpurun_out/r2_full python bench.py --steps 2 --warmup 3 --tracks 64 --track-steps 100000 --no-cpu-baseline > gpurun_out/r2_ncu_full.log 2>&1; echo "rc=$?" >> gpurun_out/r2_ncu_full.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_list.log 2>&1; echo "rc=$?" >> gpurun_out/r2_ncu_list.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_full.log 2>&1; echo "rc=$?" >> gpurun_out/r2_bench_full.log
tail -n 3 gpurun_out/r2_*.log
