#!/bin/bash
# round 2, run 14: aggregate check with its three chains on separate streams
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_verify_aggregate.py tests/test_gpu_parity.py -m gpu -x -q -k "aggregate or verify or small_order" 2>&1 | tail -3
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_c2_r2s.json 2> gpurun_out/bench_r2s.err
timeout 600 python bench.py --workload config1 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_c1_r2s.json 2>> gpurun_out/bench_r2s.err
ZKA_LANES=1 timeout 600 python bench.py --workload config1 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_c1_r2s_l1.json 2>> gpurun_out/bench_r2s.err
tail -3 gpurun_out/bench_r2s.err
python - <<'PY'
import json
for f in ('gpurun_out/bench_c2_r2s.json','gpurun_out/bench_c1_r2s.json','gpurun_out/bench_c1_r2s_l1.json'):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); v=d['verify']
            print(f, 'prove', round(d['value']), 'e2e', round(d['e2e']['value']), 'verify', round(v['value']), round(v['ms_per_step'],2), 'v_e2e', round(v['e2e']['value']), d['gpu_launches'], v['gpu_launches_per_step'])
PY
