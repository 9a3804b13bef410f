#!/bin/bash
# ncu captures of the hot kernels (run under gpurun, one GPU).  Numbers printed by bench.py under ncu are not bench values.
#   tools/ncu_capture.sh <tag> flops   every launch of one short bench run per workload with the FP64 instruction counters,
#                                      FP64 pipe activity and DRAM bytes -> gpurun_out/<tag>_flops_<workload>.csv + unit counts;
#                                      tools/ncu_flops_json.py <tag> turns them into profiles/executed_flops.json
#   tools/ncu_capture.sh <tag> full    `--set full` reports of a mid-ray slab of the three polarized stages and of the geodesic
#                                      and unpolarized kernels -> gpurun_out/<tag>_*.ncu-rep (tools/ncu_raw_summary.py)
tag=${1:-r02}
what=${2:-flops}
FP64=smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum
if [ "$what" = flops ]; then
  M=$FP64,gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum
  flops() {  # workload resolution
    BL_POL_SLAB=64 timeout 900 ncu --metrics $M --clock-control none -k regex:'pol_|geodesic|radiate' --csv --log-file gpurun_out/${tag}_flops_$1.csv \
      python bench.py --workload $1 --resolution $2 --steps 1 --warmup 0 --no-cpu-baseline --no-extras --dump-units gpurun_out/${tag}_units_$1.json \
      > gpurun_out/${tag}_flops_$1.log 2>&1
    tail -1 gpurun_out/${tag}_flops_$1.log | cut -c1-200
  }
  flops c4 256
  flops polarized_thermal 384
  flops simulation 512
  flops formula 256
  flops true_color 256
  flops render 512
else
  full() {  # name kernel-regex skip count bench-args...
    name=$1; regex=$2; skip=$3; count=$4; shift 4
    BL_POL_SLAB=64 timeout 900 ncu --set full --metrics $FP64 --clock-control none --import-source on -k regex:$regex -s $skip -c $count -f \
      -o gpurun_out/${tag}_${name} python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-extras "$@" > gpurun_out/${tag}_${name}.log 2>&1
    tail -1 gpurun_out/${tag}_${name}.log | cut -c1-200
  }
  full split pol_ 96 4 --workload c4 --resolution 512
  full unpol 'geodesic_dp|radiate_unpolarized' 0 2 --workload simulation --resolution 512
  full formula 'geodesic_dp|radiate_unpolarized' 0 2 --workload formula --resolution 384
fi
ls -la gpurun_out/${tag}_* | tail -20
