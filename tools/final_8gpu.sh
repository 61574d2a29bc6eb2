#!/bin/bash
# usage (under gpurun --gpus N): tools/final_8gpu.sh N -- the default C3 line (as the driver runs it) and the C5 line at N GPUs
N=${1:-8}
run() { tag=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@" > gpurun_out/r02_final_${tag}_${N}gpu.json 2> gpurun_out/r02_final_${tag}_${N}gpu.err || tail -5 gpurun_out/r02_final_${tag}_${N}gpu.err; }
run c3 --steps 8 --warmup 3 --no-cpu
run c5 --config c5 --steps 3 --warmup 2 --no-cpu
for f in gpurun_out/r02_final_*_${N}gpu.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], 'ms', round(d['value'],3), 'err', d['rel_err_vs_exact_u'], 'P', d['config']['partitions_per_gpu'], {k:round(v,3) for k,v in d['stage_ms'].items()}, 'e2e', d['e2e']['value'])
except Exception as e: print(sys.argv[1], 'FAILED', e)
PY
done
