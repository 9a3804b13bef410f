#!/bin/bash
# Round 2: the C4 frame on the 1.3 GB snapshot (--grid-scale 4: 308 x 256 x 512 cells, the cell gather leaves L2), live times
# and one ncu capture of a dense slab of the sampling kernel.
set -x
mkdir -p gpurun_out
timeout 900 python bench.py --workload c4 --resolution 1024 --grid-scale 4 --steps 3 --warmup 2 --no-cpu-baseline --no-extras > gpurun_out/r02aa_c4_grid4.json 2> gpurun_out/r02aa_c4_grid4.err
tail -2 gpurun_out/r02aa_c4_grid4.err | cut -c1-300
FP64=smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum
BL_POL_SLAB=64 timeout 900 ncu --set full --metrics $FP64 --clock-control none -k regex:pol_sampling -s 24 -c 1 -f -o gpurun_out/r02aa_sampling_grid4 python bench.py --workload c4 --resolution 512 --grid-scale 4 --steps 1 --warmup 0 --no-cpu-baseline --no-extras > gpurun_out/r02aa_ncu.log 2>&1
tail -2 gpurun_out/r02aa_ncu.log | cut -c1-200
