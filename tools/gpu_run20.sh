#!/bin/bash
set -x
mkdir -p gpurun_out
ZKA_TRACE=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_c2_r2t_trace.json 2> gpurun_out/bench_r2t_trace.err
grep VTRACE gpurun_out/bench_r2t_trace.err | tail -8
