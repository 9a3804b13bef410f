#!/bin/bash
# Round 2, sixteenth GPU call: geometry stage split into a sampling kernel and a frame / transport-matrix kernel.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -k "ray_ordering or pipeline_matches or cks or golden_polarized or live_reference_polarized or full_resolution or waves or slow_light or multi_device" 2>&1 | tail -8 > gpurun_out/r02p_pytest.txt
cat gpurun_out/r02p_pytest.txt
run() { tag=$1; wl=$2; res=$3; shift 3; env "$@" timeout 300 python bench.py --workload $wl --resolution $res --steps 3 --warmup 2 --no-cpu-baseline --no-extras > gpurun_out/r02p_$tag.json 2> gpurun_out/r02p_$tag.err; }
run c4_g3s5 c4 1024 BL_POL_OCC=3,0,5,5
run c4_g3s6 c4 1024 BL_POL_OCC=3,0,5,6
run c4_g4s5 c4 1024 BL_POL_OCC=4,0,5,5
run c4_g4s6 c4 1024 BL_POL_OCC=4,0,5,6
run c4_g3s4 c4 1024 BL_POL_OCC=3,0,5,4
run polth polarized_thermal 1024 A=1
