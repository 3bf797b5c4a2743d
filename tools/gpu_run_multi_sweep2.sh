#!/bin/bash
# multi-GPU: one prove call per step, finished chunks gathered while later chunks are proved — chunk size sweep; $1 = GPUs
N=${1:-2}
set -x
mkdir -p gpurun_out
for C in 4096 1408 928; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952${C:0:1} bench.py --gpus $N --steps 4 --warmup 3 --no-cpu --gather-chunk $C > gpurun_out/bench_c3_${N}gpu_c${C}_r2n.json 2> gpurun_out/bench_c3_${N}gpu_c${C}_r2n.err
  grep -v "NCCL INFO" gpurun_out/bench_c3_${N}gpu_c${C}_r2n.err | grep -i "error\|assert" | head -5
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_c3_*gpu_c*_r2n.json')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); g=d.get('gather',{})
            print(f, round(d['value']), round(d['ms_per_step'],2), 'verify', round(d['verify']['value']), g.get('exposed_ms_per_step'), g.get('group_rows'), g.get('own_rows_roundtrip'))
PY
