#!/bin/bash
# Round 2, second GPU call: occupancy sweep of the three polarized stages, sparsified coupling, ncu capture per stage.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "polarized" 2>&1 | tail -5
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --workload polarized --resolution 1024 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02b_$tag.json 2> gpurun_out/r02b_$tag.err; }
run base BL_POL_OCC=2,4,4
run g3 BL_POL_OCC=3,4,4
run g4 BL_POL_OCC=4,4,4
run c3 BL_POL_OCC=2,3,4
run c5 BL_POL_OCC=2,5,4
run c6 BL_POL_OCC=2,6,4
run t3 BL_POL_OCC=2,4,3
run t5 BL_POL_OCC=2,4,5
run t6 BL_POL_OCC=2,4,6
run fused BL_POL_FUSED=1
FP64=smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum
BL_POL_SLAB=64 timeout 900 ncu --set full --metrics $FP64 --clock-control none --import-source on -k regex:pol_ -s 51 -c 6 -f \
  -o gpurun_out/r02b_split python bench.py --workload polarized --resolution 384 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r02b_ncu.log 2>&1
tail -3 gpurun_out/r02b_ncu.log | cut -c1-300
# flop totals of one whole pass (all launches), for flop-per-sample accounting
BL_POL_SLAB=64 timeout 900 ncu --metrics $FP64,gpu__time_duration.sum --clock-control none -k regex:'pol_|geodesic' --csv --log-file gpurun_out/r02b_flops.csv \
  python bench.py --workload polarized --resolution 256 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r02b_flops.log 2>&1
tail -2 gpurun_out/r02b_flops.log | cut -c1-600
ls -la gpurun_out | tail -8
