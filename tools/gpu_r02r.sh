#!/bin/bash
# Round 2, eighteenth GPU call: factored Stokes coupling in the transfer stage; integrator occupancy by rays per thread.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -k "polarized or cks or full_resolution or adaptive_two_levels or slow_light or iharm3d" 2>&1 | tail -8 > gpurun_out/r02r_pytest.txt
cat gpurun_out/r02r_pytest.txt
run() { tag=$1; wl=$2; res=$3; shift 3; env "$@" timeout 300 python bench.py --workload $wl --resolution $res --steps 3 --warmup 2 --no-cpu-baseline --no-extras > gpurun_out/r02r_$tag.json 2> gpurun_out/r02r_$tag.err; }
run c4_t5 c4 1024 A=1
run c4_t6 c4 1024 BL_POL_OCC=3,0,6,5
run c4_t4 c4 1024 BL_POL_OCC=3,0,4,5
run polth polarized_thermal 1024 A=1
run formula formula 512 A=1
run adaptive adaptive 512 A=1
