#!/bin/bash
# Round 2, third GPU call: the rewritten bench (main workload + extras), the flop-count capture, polarized parity again.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "polarized" 2>&1 | tail -5
timeout 900 python bench.py --resolution 1024 --steps 2 --warmup 1 > gpurun_out/r02c_bench_1024.json 2> gpurun_out/r02c_bench_1024.err
tail -5 gpurun_out/r02c_bench_1024.err
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r02c_ref.json 2> gpurun_out/r02c_ref.err
tail -3 gpurun_out/r02c_ref.err
bash tools/ncu_capture.sh r02c flops
