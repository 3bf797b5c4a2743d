#!/bin/bash
# round 2, run 18: occupancy variants of the commit / MSM kernels (__launch_bounds__ min blocks 5 and 6), memcheck of smoke()
set -x
mkdir -p gpurun_out
for mb in 5 6; do
  if [ -f zkp_ecdsa_b200/libzkattest_mb$mb.so ]; then
    ZKA_LIB=zkp_ecdsa_b200/libzkattest_mb$mb.so timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_c2_r2v_mb$mb.json 2>> gpurun_out/bench_r2v.err
  fi
done
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_c2_r2v_base.json 2>> gpurun_out/bench_r2v.err
timeout 900 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 > gpurun_out/compute_sanitizer_memcheck_r2z.txt; cat gpurun_out/compute_sanitizer_memcheck_r2z.txt
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_c2_r2v_*.json')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); v=d['verify']
            print(f, 'prove', round(d['value']), 'verify', round(v['value']), round(d['roofline']['frac'],3))
            print('   P', [(k, x['ms_per_step']) for k, x in list(d['kernels'].items())[:6]])
            print('   V', [(k, x['ms_per_step']) for k, x in list(v['kernels'].items())[:3]])
PY
