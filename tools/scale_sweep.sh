#!/bin/bash
# usage (under gpurun --gpus N): tools/scale_sweep.sh N  -- C3 partition sweep + C5 lines at N GPUs, JSON lines to gpurun_out/
N=${1:-8}
run() { tag=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@" > gpurun_out/r02_sweep_${tag}_${N}gpu.json 2> gpurun_out/r02_sweep_${tag}_${N}gpu.err || tail -5 gpurun_out/r02_sweep_${tag}_${N}gpu.err; }
for P in 148 222 296; do run c3_P$P --steps 5 --warmup 3 --no-e2e --no-cpu --partitions $P; done
run c5_P64 --config c5 --steps 3 --warmup 2 --no-cpu --partitions 64
run c5_P128 --config c5 --steps 3 --warmup 2 --no-cpu --partitions 128
for f in gpurun_out/r02_sweep_*_${N}gpu.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], 'ms', round(d['value'],3), 'err', d['rel_err_vs_exact_u'], 'P', d['config']['partitions_per_gpu'], {k:round(v,3) for k,v in d['stage_ms'].items()}, 'rf', round(d['roofline']['frac'],3))
except Exception as e: print(sys.argv[1], 'FAILED', e)
PY
done
