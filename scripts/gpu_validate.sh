# Final validation of a tree on one B200: GPU suite, smoke, default bench line (+ reference arm), ncu launch
# list of the same command, ncu --set full captures of the dominant kernels, per-config timings.
mkdir -p gpurun_out
T=${1:-val}
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_smoke.log
timeout 600 python bench.py > gpurun_out/${T}_bench.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_bench.log
timeout 600 python bench.py --impl reference > gpurun_out/${T}_bench_ref.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_bench_ref.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_bench.log 2>&1
NCU_SKIP=6 timeout 600 bash scripts/ncu_capture.sh ${T}_ctcrw "ctcrw_fwd|ctcrw_bwd" 2 python bench.py --workload tracks --steps 2 --warmup 3 --no-cpu-baseline
NCU_SKIP=2 timeout 600 bash scripts/ncu_capture.sh ${T}_ou "sde_stream" 1 python scripts/ou_run.py 4096 25000 4
rm -f gpurun_out/${T}_ctcrw.source.csv gpurun_out/${T}_ou.source.csv gpurun_out/${T}_ou.raw.csv
grep -v Warn gpurun_out/${T}_pytest.log | tail -4; tail -3 gpurun_out/${T}_smoke.log; tail -n 2 gpurun_out/${T}_bench.log | cut -c1-400
