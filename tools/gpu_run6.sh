#!/bin/bash
# round 2, run 6: full GPU test-suite + default bench with the split tape upload / exact-length row copies
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_c2_r2g.json 2> gpurun_out/bench_r2g.err
timeout 600 python bench.py --workload config1 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_c1_r2g.json 2>> gpurun_out/bench_r2g.err
tail -5 gpurun_out/bench_r2g.err
python - <<'PY'
import json
for f in ('gpurun_out/bench_c2_r2g.json','gpurun_out/bench_c1_r2g.json'):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f, round(d['value']), round(d['e2e']['value']), round(d['verify']['value']), d['verify'].get('e2e'))
PY
