#!/bin/bash
# Round 2, 2-GPU call: sharded adaptive run with device-resident gather and assembly.
set -x
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0 timeout 600 python -m pytest tests -m gpu -q -x -k "adaptive_sharded" 2>&1 | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --workload adaptive --steps 3 --warmup 2 > gpurun_out/r02y_adaptive_n2.json 2> gpurun_out/r02y_adaptive_n2.err
tail -2 gpurun_out/r02y_adaptive_n2.err | cut -c1-300
cut -c1-900 gpurun_out/r02y_adaptive_n2.json
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --workload adaptive --steps 3 --warmup 2 > gpurun_out/r02y_adaptive_n1.json 2> gpurun_out/r02y_adaptive_n1.err
cut -c1-900 gpurun_out/r02y_adaptive_n1.json
