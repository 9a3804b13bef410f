#!/bin/bash
# Round 2, eighth GPU call (2 GPUs): the bench through torchrun at N = 2 (main workload at 2048^2 + all extras).
set -x
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --resolution 2048 --steps 2 --warmup 1 > gpurun_out/r02h_bench_n2.json 2> gpurun_out/r02h_bench_n2.err
tail -5 gpurun_out/r02h_bench_n2.err | cut -c1-300
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --resolution 2048 --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/r02h_bench_n1.json 2> gpurun_out/r02h_bench_n1.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/r02h_ref_n2.json 2> gpurun_out/r02h_ref_n2.err
cat gpurun_out/r02h_ref_n2.json | cut -c1-300
