#!/bin/bash
set -x
mkdir -p gpurun_out
free -g | head -2; nproc
ZKA_LANES=3 timeout 900 python bench.py --batch 32768 --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_c2_r2f_b32768.json 2>> gpurun_out/bench_r2f.err
ZKA_LANES=3 ZKA_TRACE=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_c2_r2f_trace.json 2> gpurun_out/bench_r2f_trace.err
tail -3 gpurun_out/bench_r2f.err
grep -c TRACE gpurun_out/bench_r2f_trace.err
