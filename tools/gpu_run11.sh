#!/bin/bash
# round 2, run 11: verify host path with prefetched inputs (timeline trace + bench), aggregate check after the tail changes
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_verify_aggregate.py tests/test_gpu_parity.py tests/test_war256.py -m gpu -x -q -k "aggregate or verify or pipeline or war" 2>&1 | tail -4
for i in 1 2; do
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_c2_r2l_$i.json 2>> gpurun_out/bench_r2l.err
done
for cfg in "6 2048" "4 4096"; do
  set -- $cfg
  ZKA_LANES=$1 ZKA_HOST_CHUNK=$2 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_c2_r2l_l$1_h$2.json 2>> gpurun_out/bench_r2l.err
done
ZKA_TRACE=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_c2_r2l_trace.json 2> gpurun_out/bench_r2l_trace.err
grep -c TRACE gpurun_out/bench_r2l_trace.err
timeout 600 python bench.py --workload config1 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_c1_r2l.json 2>> gpurun_out/bench_r2l.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_c?_r2l*.json')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); v=d['verify']
            print(f, 'prove', round(d['value']), 'e2e', round(d['e2e']['value']), 'verify', round(v['value']), 'v_e2e', round(v['e2e']['value']), (v.get('roofline') or {}).get('frac'))
            print('   V', [(k, x['ms_per_step']) for k, x in list(v['kernels'].items())[:8]])
PY
