#!/bin/bash
# Round 2, tenth GPU call: coefficient-stage unroll variants.
set -x
mkdir -p gpurun_out
run() { tag=$1; wl=$2; res=$3; shift 3; env "$@" timeout 300 python bench.py --workload $wl --resolution $res --steps 3 --warmup 2 --no-cpu-baseline --no-extras > gpurun_out/r02j_$tag.json 2> gpurun_out/r02j_$tag.err; }
run c4_base c4 1024 BL_POL_OCC=3,4,5
run c4_u2_m3 c4 1024 BL_POL_OCC=3,23,5
run c4_u2_m4 c4 1024 BL_POL_OCC=3,24,5
run c4_m3 c4 1024 BL_POL_OCC=3,3,5
run polth_base polarized_thermal 1024 BL_POL_OCC=3,4,5
run polth_m5 polarized_thermal 1024 BL_POL_OCC=3,5,5
