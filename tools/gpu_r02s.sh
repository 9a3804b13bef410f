#!/bin/bash
# Round 2, nineteenth GPU call: coefficient-stage unrolling over frequencies, record prefetch into L1.
set -x
mkdir -p gpurun_out
run() { tag=$1; wl=$2; res=$3; shift 3; env "$@" timeout 300 python bench.py --workload $wl --resolution $res --steps 3 --warmup 2 --no-cpu-baseline --no-extras > gpurun_out/r02s_$tag.json 2> gpurun_out/r02s_$tag.err; }
run c4_u1 c4 1024 A=1
run c4_u2 c4 1024 BL_POL_UNROLL=2
run c4_u4 c4 1024 BL_POL_UNROLL=4
run c4_u4_c5 c4 1024 BL_POL_UNROLL=4 BL_POL_OCC=3,5,5,5
run c4_u4_c4 c4 1024 BL_POL_UNROLL=4 BL_POL_OCC=3,4,5,5
run c4_u2_c5 c4 1024 BL_POL_UNROLL=2 BL_POL_OCC=3,5,5,5
run c4_pf11 c4 1024 BL_RAD_PREFETCH=11
run c4_pf0 c4 1024 BL_RAD_PREFETCH=0
run sim_pf11 simulation 1024 BL_RAD_PREFETCH=11
run sim_pf12 simulation 1024 BL_RAD_PREFETCH=12
run sim_pf2 simulation 1024 A=1
