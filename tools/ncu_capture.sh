#!/bin/bash
# One `ncu --set full` capture per hot kernel (run under gpurun, one GPU): geodesic + unpolarized at 512^2,
# polarized thermal at 384^2, polarized kappa (4 frequencies) at 256^2.  Reports land in gpurun_out/<tag>_*.ncu-rep;
# summarise here with tools/ncu_raw_summary.py.  Numbers printed by bench.py under ncu are not bench values.
tag=${1:-r01k}
FP64=smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum
run() {  # name kernel-regex bench-args...
  name=$1; regex=$2; shift 2
  timeout 600 ncu --set full --metrics $FP64 --clock-control none --import-source on -k regex:$regex -c ${COUNT:-1} -f \
    -o gpurun_out/${tag}_${name} python bench.py --steps 1 --warmup 0 --no-cpu-baseline "$@" > gpurun_out/${tag}_${name}.log 2>&1
  tail -2 gpurun_out/${tag}_${name}.log | cut -c1-200
}
COUNT=2 run unpol 'geodesic_dp|radiate_unpolarized' --resolution 512
run polth radiate_polarized --workload polarized_thermal --resolution 384
run polk4 radiate_polarized --workload polarized --resolution 256
ls -la gpurun_out/${tag}_*.ncu-rep
