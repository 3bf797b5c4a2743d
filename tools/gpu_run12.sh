#!/bin/bash
# round 2, run 12: final-build regression + bench lines + ncu evidence (launch list with pipe counters, --set full of the
# dominant kernels of both legs) + memcheck of smoke()
set -x
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c2_r2m.json 2> gpurun_out/bench_r2m.err
timeout 600 python bench.py --workload config1 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_c1_r2m.json 2>> gpurun_out/bench_r2m.err
timeout 600 python bench.py --workload config3 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_c3_r2m.json 2>> gpurun_out/bench_r2m.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r2m.json 2>> gpurun_out/bench_r2m.err
timeout 600 python tools/bench_war256.py 8192 256 3 > gpurun_out/bench_war256_r2m.json 2>> gpurun_out/bench_r2m.err
tail -3 gpurun_out/bench_r2m.err
M=gpu__time_duration.sum,sm__inst_executed_pipe_fmaheavy.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__cycles_elapsed.max
ZKA_LANES=1 timeout 900 ncu --profile-from-start off --metrics $M --clock-control none --kernel-name-base demangled --csv --log-file gpurun_out/launches_r2m_config2_pipes.csv python tools/profile_step.py 2>&1 | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -c 600 --csv --log-file gpurun_out/launches_r2m_bench_config2.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu_r2m.log 2>&1
for k in "AggBucketTask<zk::AggTomSrc>" TomCommitH PhaseAAndRPoint AggTorsionPart; do
  n=$(echo $k | tr -cd 'A-Za-z')
  ZKA_LANES=1 timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none --kernel-name-base demangled -k "regex:$k" -c 1 -f -o /tmp/ncu_$n python tools/profile_step.py 2>&1 | tail -1
  ncu -i /tmp/ncu_$n.ncu-rep --page details --csv > gpurun_out/ncu_r2m_$n.details.csv 2>/dev/null
  ncu -i /tmp/ncu_$n.ncu-rep --page raw --csv > gpurun_out/ncu_r2m_$n.raw.csv 2>/dev/null
done
ncu -i /tmp/ncu_AggBucketTaskzkAggTomSrc.ncu-rep --page source --csv 2>/dev/null | head -c 4000000 > gpurun_out/ncu_r2m_AggBucket.source.csv
timeout 900 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 > gpurun_out/compute_sanitizer_memcheck_r2m.txt; cat gpurun_out/compute_sanitizer_memcheck_r2m.txt
du -sh gpurun_out
python - <<'PY'
import json
for f in ('gpurun_out/bench_c2_r2m.json','gpurun_out/bench_c1_r2m.json','gpurun_out/bench_c3_r2m.json'):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); v=d['verify']
            print(f, 'prove', round(d['value']), 'e2e', round(d['e2e']['value']), 'verify', round(v['value']), 'v_e2e', round(v['e2e']['value']), d['roofline']['frac'], d['roofline']['whole_step']['frac'], (v.get('roofline') or {}).get('frac'), d.get('cpu_baseline'))
for f in ('gpurun_out/bench_ref_r2m.json','gpurun_out/bench_war256_r2m.json'):
    print(open(f).read()[:600])
PY
