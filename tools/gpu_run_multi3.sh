#!/bin/bash
# multi-GPU: pipelined gather (default) vs serial; $1 = GPUs
N=${1:-2}
set -x
mkdir -p gpurun_out
for M in pipeline serial; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu --gather-mode $M > gpurun_out/bench_c3_${N}gpu_${M}_r2p.json 2> gpurun_out/bench_c3_${N}gpu_${M}_r2p.err
  grep -v "NCCL INFO" gpurun_out/bench_c3_${N}gpu_${M}_r2p.err | grep -i "error\|assert" | head -5
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_c3_*gpu_*_r2p.json')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); g=d.get('gather',{})
            print(f, round(d['value']), round(d['ms_per_step'],2), 'verify', round(d['verify']['value']), g.get('exposed_ms_per_step'), g.get('mode'), g.get('own_rows_roundtrip'), g.get('checksums_match_all_ranks'))
PY
