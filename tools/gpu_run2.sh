#!/bin/bash
# round-2 second GPU call: new parity tests, lanes/chunk sweep, captures with demangled kernel names
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for cfg in "2 4096" "1 8192" "2 2048" "3 2048" "3 4096" "4 2048"; do
  set -- $cfg
  ZKA_LANES=$1 ZKA_CHUNK=$2 ZKA_HOST_CHUNK=$2 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_c2_r2b_l$1_c$2.json 2>> gpurun_out/bench_r2b.err
done
for cfg in "1 8192" "2 512" "3 512" "4 256"; do
  set -- $cfg
  ZKA_LANES=$1 ZKA_CHUNK=$2 ZKA_HOST_CHUNK=$2 timeout 600 python bench.py --workload config1 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_c1_r2b_l$1_c$2.json 2>> gpurun_out/bench_r2b.err
done
ZKA_LIB=zkp_ecdsa_b200/libzkattest_inl.so ZKA_LANES=2 ZKA_CHUNK=4096 ZKA_HOST_CHUNK=4096 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_c2_r2b_inl_l2_c4096.json 2>> gpurun_out/bench_r2b.err
ZKA_LIB=zkp_ecdsa_b200/libzkattest_inl.so ZKA_LANES=1 ZKA_CHUNK=8192 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_c2_r2b_inl_l1_c8192.json 2>> gpurun_out/bench_r2b.err
tail -5 gpurun_out/bench_r2b.err
ls -la gpurun_out | tail -20
