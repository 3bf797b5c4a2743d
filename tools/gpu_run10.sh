#!/bin/bash
# round 2, run 10: e2e sensitivity to lanes / host chunk (prove and verify host-buffer legs)
set -x
mkdir -p gpurun_out
for cfg in "3 2048" "4 2048" "6 2048" "6 1024" "3 4096" "4 4096"; do
  set -- $cfg
  ZKA_LANES=$1 ZKA_HOST_CHUNK=$2 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_c2_r2k_l$1_h$2.json 2>> gpurun_out/bench_r2k.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_c2_r2k_*.json')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); v=d['verify']
            print(f, 'prove', round(d['value']), 'e2e', round(d['e2e']['value']), 'verify', round(v['value']), 'v_e2e', round(v['e2e']['value']))
PY
