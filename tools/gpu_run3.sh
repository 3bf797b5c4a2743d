#!/bin/bash
# round-2 third GPU call: tapered dynamic schedule sweep, K-sample verifier test, captures (kept under 64 MiB)
set -x
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for cfg in "3 4096 2048" "2 4096 2048" "3 4096 1024" "4 4096 2048" "3 8192 4096" "4 2048 1024"; do
  set -- $cfg
  ZKA_LANES=$1 ZKA_CHUNK=$2 ZKA_HOST_CHUNK=$3 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_c2_r2c_l$1_c$2_h$3.json 2>> gpurun_out/bench_r2c.err
done
for cfg in "3 4096 2048" "2 4096 2048" "4 4096 2048"; do
  set -- $cfg
  ZKA_LANES=$1 ZKA_CHUNK=$2 ZKA_HOST_CHUNK=$3 timeout 600 python bench.py --workload config1 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_c1_r2c_l$1_c$2_h$3.json 2>> gpurun_out/bench_r2c.err
done
timeout 600 python bench.py --workload config3 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_c3_r2c.json 2>> gpurun_out/bench_r2c.err
tail -5 gpurun_out/bench_r2c.err
SEC="--section SpeedOfLight --section ComputeWorkloadAnalysis --section MemoryWorkloadAnalysis --section Occupancy --section LaunchStats --section WarpStateStats --section SchedulerStats --section InstructionStats"
# .ncu-rep files are ~14-20 MB each and gpurun_out/ is capped at 64 MiB: export the text pages on the box, drop the reports
for k in MsmTomWindowBoth PhaseAAndRPoint TomCommitH TomNormTask VValidate; do
  ZKA_LANES=1 ZKA_CHUNK=8192 timeout 400 ncu --profile-from-start off $SEC --clock-control none --kernel-name-base demangled -k regex:$k -c 1 -f -o /tmp/ncu_$k python tools/profile_step.py 2>&1 | tail -1
  ncu -i /tmp/ncu_$k.ncu-rep --page details --csv > gpurun_out/ncu_r2c_$k.details.csv 2>/dev/null
  ncu -i /tmp/ncu_$k.ncu-rep --page raw --csv > gpurun_out/ncu_r2c_$k.raw.csv 2>/dev/null
done
ZKA_LANES=1 ZKA_CHUNK=8192 timeout 400 ncu --profile-from-start off --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:MsmTomWindowBoth -c 1 -f -o /tmp/ncu_full_msm python tools/profile_step.py 2>&1 | tail -1
ncu -i /tmp/ncu_full_msm.ncu-rep --page details --csv > gpurun_out/ncu_r2c_full_MsmTomWindowBoth.details.csv 2>/dev/null
ncu -i /tmp/ncu_full_msm.ncu-rep --page raw --csv > gpurun_out/ncu_r2c_full_MsmTomWindowBoth.raw.csv 2>/dev/null
ncu -i /tmp/ncu_full_msm.ncu-rep --page source --csv 2>/dev/null | head -c 6000000 > gpurun_out/ncu_r2c_full_MsmTomWindowBoth.source.csv
du -sh gpurun_out; ls -la gpurun_out | tail -25
