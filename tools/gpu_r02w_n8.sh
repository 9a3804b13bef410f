#!/bin/bash
# Round 2, 8-GPU call: the default bench (4096^2 C4 frame, strong scaling, all extras) on 8 ranks, the main workload on 4,
# and the C++ multi-device drop-in driver on the C4 frame (2048^2) with 8 devices against one.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv | head -10
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02w_bench_n8.json 2> gpurun_out/r02w_bench_n8.err
tail -3 gpurun_out/r02w_bench_n8.err | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 4 --steps 3 --warmup 3 --no-extras > gpurun_out/r02w_bench_n4.json 2> gpurun_out/r02w_bench_n4.err
tail -3 gpurun_out/r02w_bench_n4.err | cut -c1-300
timeout 600 python tools/dropin_multi.py 2048 > gpurun_out/r02w_dropin.txt 2>&1
tail -5 gpurun_out/r02w_dropin.txt
