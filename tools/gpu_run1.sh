#!/bin/bash
# round-2 first GPU call: sanity tests, config2 bench line, per-launch pipe counters, full captures
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c2_r2a.json 2> gpurun_out/bench_c2_r2a.err; tail -3 gpurun_out/bench_c2_r2a.err
M=gpu__time_duration.sum,sm__inst_executed_pipe_fmaheavy.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__cycles_elapsed.max
timeout 600 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/launches_r2a_config2_pipes.csv python tools/profile_step.py 2>&1 | tail -2
for k in MsmTomWindowBoth PhaseAAndRPoint TomCommitH TomNormTask; do
  timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o gpurun_out/ncu_r2a_$k python tools/profile_step.py 2>&1 | tail -2
done
ls -la gpurun_out
