mkdir -p gpurun_out
T=${1:-scale}
run() { # N workload
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29600 + $1 + ${3:-0})) bench.py --gpus $1 --workload $2 --steps 20 --warmup 3 > gpurun_out/${T}_$2_$1.log 2>&1
  python - gpurun_out/${T}_$2_$1.log $1 $2 <<'PY'
import json,sys
ls=[x for x in open(sys.argv[1]) if x.startswith('{')]
if not ls: print(sys.argv[2:], 'FAILED', open(sys.argv[1]).read()[-600:])
else:
    d=json.loads(ls[-1]); print(sys.argv[3], 'N=%s'%sys.argv[2], 'ms/step %.3f'%d['ms_per_step'], 'value %.4g'%d['value'], 'e2e %.4g'%d['e2e']['value'])
PY
}
run 8 tracks
run 8 single 20
run 4 tracks
