#!/bin/bash
# round 2, run 9: full GPU suite after the proof-group refactor, war256 build (tests + throughput), smoke, PCIe bandwidth
set -x
mkdir -p gpurun_out
python tools/pcie_bw.py > gpurun_out/pcie_bw.json 2>/dev/null; cat gpurun_out/pcie_bw.json
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python tools/bench_war256.py 8192 256 3 > gpurun_out/bench_war256_r2j.json 2>> gpurun_out/bench_r2j.err; cat gpurun_out/bench_war256_r2j.json
ZKA_BENCH_GROUP=tomEdwards256 timeout 600 python tools/bench_war256.py 8192 256 3 > gpurun_out/bench_tom_same_tool_r2j.json 2>> gpurun_out/bench_r2j.err; cat gpurun_out/bench_tom_same_tool_r2j.json
for i in 1 2; do
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_c2_r2j_$i.json 2>> gpurun_out/bench_r2j.err
done
tail -5 gpurun_out/bench_r2j.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_c2_r2j_*.json')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); v=d['verify']
            print(f, 'prove', round(d['value']), 'e2e', round(d['e2e']['value']), 'verify', round(v['value']), 'v_e2e', round(v['e2e']['value']), v.get('all_accepted'), d['roofline']['frac'], (v.get('roofline') or {}).get('frac'))
PY
