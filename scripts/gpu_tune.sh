# usage: bash scripts/gpu_tune.sh <tag> <suffix> [<suffix> ...] : bench every tuning build, print ms and kernel times
mkdir -p gpurun_out
T=$1; shift
for s in "$@"; do
  SSDE_LIB_SUFFIX=_$s timeout 300 python bench.py --workload tracks --no-cpu-baseline --steps 20 > gpurun_out/${T}_$s.log 2>&1
  python - "$s" gpurun_out/${T}_$s.log <<'PY'
import json,sys
ls=[x for x in open(sys.argv[2]) if x.startswith('{')]
if not ls: print(sys.argv[1], 'FAILED', open(sys.argv[2]).read()[-400:])
else:
    d=json.loads(ls[-1]); k=d['roofline']['kernels_ms']; li=d.get('launch_info')
    print(f"{sys.argv[1]:12s} ms/step {d['ms_per_step']:.3f}  fwd {k['ctcrw_fwd']:.3f}  bwd {k['ctcrw_bwd']:.3f}  nllk {d['nllk']:.10g} launch {li}")
PY
done
