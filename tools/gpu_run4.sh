#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for cfg in "3 4096 2048" "2 4096 2048" "3 4096 1024" "4 4096 2048" "4 2048 1024" "3 2048 2048"; do
  set -- $cfg
  ZKA_LANES=$1 ZKA_CHUNK=$2 ZKA_HOST_CHUNK=$3 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_c2_r2d_l$1_c$2_h$3.json 2>> gpurun_out/bench_r2d.err
done
for cfg in "3 4096 2048" "2 4096 2048" "4 4096 2048"; do
  set -- $cfg
  ZKA_LANES=$1 ZKA_CHUNK=$2 ZKA_HOST_CHUNK=$3 timeout 600 python bench.py --workload config1 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_c1_r2d_l$1_c$2_h$3.json 2>> gpurun_out/bench_r2d.err
done
ZKA_LIB=zkp_ecdsa_b200/libzkattest_msm4.so timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_c2_r2d_msm4.json 2>> gpurun_out/bench_r2d.err
timeout 600 python bench.py --workload config3 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_c3_r2d.json 2>> gpurun_out/bench_r2d.err
tail -5 gpurun_out/bench_r2d.err
du -sh gpurun_out
