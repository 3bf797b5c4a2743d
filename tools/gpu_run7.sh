#!/bin/bash
# round 2, run 7: chunk-wide aggregate verify check — tests, A/B against the per-proof path, window sweep; tape split A/B
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_verify_aggregate.py tests/test_gpu_parity.py -m gpu -x -q -k "aggregate or verify" 2>&1 | tail -5
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_c2_r2h_$name.json 2>> gpurun_out/bench_r2h.err
}
run agg1 ZKA_AGG=1
run agg0 ZKA_AGG=0
for c in 12 13 15 16; do run aggc$c ZKA_AGG_C=$c; done
run split0 ZKA_TAPE_SPLIT=0
env ZKA_AGG=1 timeout 600 python bench.py --workload config1 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_c1_r2h_agg1.json 2>> gpurun_out/bench_r2h.err
env ZKA_AGG=1 ZKA_TAPE_SPLIT=0 timeout 600 python bench.py --workload config1 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_c1_r2h_split0.json 2>> gpurun_out/bench_r2h.err
tail -5 gpurun_out/bench_r2h.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_c?_r2h_*.json')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); v=d['verify']
            print(f, 'prove', round(d['value']), 'e2e', round(d['e2e']['value']), 'verify', round(v['value']), 'v_e2e', round(v['e2e']['value']), v.get('all_accepted'))
            top=list(v.get('kernels',{}).items())[:9] if isinstance(v.get('kernels'),dict) else v.get('kernels')
            print('   ', [(k, x['ms_per_step']) for k, x in top] if top else None)
PY
