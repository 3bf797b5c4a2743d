#!/bin/bash
# multi-GPU bench under torchrun: $1 = number of GPUs
N=${1:-2}
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv | head -10
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 4 --warmup 3 > gpurun_out/bench_c3_${N}gpu_r2z.json 2> gpurun_out/bench_c3_${N}gpu_r2z.err
tail -c 1500 gpurun_out/bench_c3_${N}gpu_r2z.err | tail -15
grep -c "NCCL INFO" gpurun_out/bench_c3_${N}gpu_r2z.err
head -c 600 gpurun_out/bench_c3_${N}gpu_r2z.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/bench_ref_${N}gpu_r2z.json 2>> gpurun_out/bench_c3_${N}gpu_r2z.err
head -c 400 gpurun_out/bench_ref_${N}gpu_r2z.json
