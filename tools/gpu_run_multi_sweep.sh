#!/bin/bash
# multi-GPU: sub-batch groups of the overlapped all-gather (1 = gather after the whole batch), $1 = GPUs
N=${1:-2}
set -x
mkdir -p gpurun_out
for G in 1 2 3; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$G bench.py --gpus $N --steps 4 --warmup 3 --no-cpu --gather-groups $G > gpurun_out/bench_c3_${N}gpu_g${G}_r2m.json 2> gpurun_out/bench_c3_${N}gpu_g${G}_r2m.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_c3_*gpu_g*_r2m.json')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f, round(d['value']), round(d['ms_per_step'],2), 'verify', round(d['verify']['value']), d.get('gather',{}).get('exposed_ms_per_step'), d.get('gather',{}).get('group_rows'))
PY
