mkdir -p gpurun_out
K='regex:ctcrw_fwd|ctcrw_bwd'
SSDE_LIB_SUFFIX=_pp1 timeout 900 ncu --set full --clock-control none --import-source on -k "$K" --launch-skip 6 --launch-count 2 -f -o gpurun_out/r38_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r38_ncu_full.log 2>&1; echo "rc=$?"
