#!/bin/bash
# round 2, run 13: final tree — full GPU suite, smoke, default bench lines (config2 with the CPU baseline leg, config1, config3)
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_c2_r2z.json 2> gpurun_out/bench_r2z.err
timeout 600 python bench.py --workload config1 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_c1_r2z.json 2>> gpurun_out/bench_r2z.err
timeout 600 python bench.py --workload config3 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_c3_r2z.json 2>> gpurun_out/bench_r2z.err
tail -3 gpurun_out/bench_r2z.err
python - <<'PY'
import json
for f in ('gpurun_out/bench_c2_r2z.json', 'gpurun_out/bench_c1_r2z.json', 'gpurun_out/bench_c3_r2z.json'):
    for l in open(f):
        if l.startswith('{'):
            d = json.loads(l); v = d['verify']
            print(f, 'prove', round(d['value']), 'e2e', round(d['e2e']['value']), 'verify', round(v['value']), 'v_e2e', round(v['e2e']['value']))
PY
