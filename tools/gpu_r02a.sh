#!/bin/bash
# Round 2, first GPU call: parity of the three-stage polarized pipeline, A/B against the fused kernel, ncu launch list.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "polarized or waves or adaptive_drop_in or full_resolution or formats" 2>&1 | tail -15
for res in 512 1024; do
  timeout 300 python bench.py --workload polarized --resolution $res --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02a_polk4_split_$res.json 2> gpurun_out/r02a_polk4_split_$res.err
  BL_POL_FUSED=1 timeout 300 python bench.py --workload polarized --resolution $res --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02a_polk4_fused_$res.json 2> gpurun_out/r02a_polk4_fused_$res.err
done
timeout 300 python bench.py --workload polarized_thermal --resolution 1024 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02a_polth_split.json 2> gpurun_out/r02a_polth_split.err
BL_POL_FUSED=1 timeout 300 python bench.py --workload polarized_thermal --resolution 1024 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02a_polth_fused.json 2> gpurun_out/r02a_polth_fused.err
for slab in 32 128 256; do
  BL_POL_SLAB=$slab timeout 300 python bench.py --workload polarized --resolution 1024 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02a_polk4_slab$slab.json 2> gpurun_out/r02a_polk4_slab$slab.err
done
FP64=smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum
# one mid-ray slab of each stage under --set full (384^2: 27 slabs x 3 launches per pass; skip the warm-up pass and the first slabs)
timeout 900 ncu --set full --metrics $FP64 --clock-control none --import-source on -k regex:'pol_(geometry|coefficient|transfer)' -s 60 -c 3 -f \
  -o gpurun_out/r02a_split python bench.py --workload polarized --resolution 384 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r02a_ncu.log 2>&1
tail -3 gpurun_out/r02a_ncu.log | cut -c1-300
ls -la gpurun_out | tail -20
