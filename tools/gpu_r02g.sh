#!/bin/bash
# Round 2, seventh GPU call: full ncu reports of the current kernels (source-level stalls) + bench of the hand-reduced connection.
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x -k "polarized" 2>&1 | tail -3
timeout 300 python bench.py --resolution 1024 --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r02g_c4_1024.json 2> gpurun_out/r02g_c4_1024.err
bash tools/ncu_capture.sh r02g full
