#!/bin/bash
# round 2, run 8: aggregate verify check with the spread top window + device-chosen key-table width
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_verify_aggregate.py tests/test_gpu_parity.py -m gpu -x -q -k "aggregate or verify or few_distinct or bit_exact" 2>&1 | tail -5
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_c2_r2i_$name.json 2>> gpurun_out/bench_r2i.err
}
run default ZKA_AGG=1
for c in 12 13 14 15 16; do run aggc$c ZKA_AGG_C=$c; done
env ZKA_AGG=1 timeout 600 python bench.py --workload config1 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_c1_r2i_default.json 2>> gpurun_out/bench_r2i.err
tail -5 gpurun_out/bench_r2i.err
if [ -f zkp_ecdsa_b200/libzkattest_war256.so ]; then
  timeout 900 python -m pytest tests/test_war256.py -m gpu -x -q 2>&1 | tail -5
  timeout 600 python tools/bench_war256.py 8192 256 3 > gpurun_out/bench_war256_r2i.json 2>> gpurun_out/bench_r2i.err
  cat gpurun_out/bench_war256_r2i.json
fi
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_c?_r2i_*.json')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); v=d['verify']
            print(f, 'prove', round(d['value']), 'e2e', round(d['e2e']['value']), 'verify', round(v['value']), 'v_e2e', round(v['e2e']['value']), v.get('all_accepted'), (v.get('roofline') or {}).get('frac'))
            top=list(v.get('kernels',{}).items())[:9]
            print('   V', [(k, x['ms_per_step']) for k, x in top])
            top=list(d.get('kernels',{}).items())[:9]
            print('   P', [(k, x['ms_per_step']) for k, x in top])
PY
