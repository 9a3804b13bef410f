#!/bin/bash
# Round 2, fifteenth GPU call: rays sorted by length for the radiation kernels (bitwise A/B, then timing), coefficient occupancy 7 / 8.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -k "ray_ordering or pipeline_matches or cks or golden_polarized or golden_unpolarized or waves or adaptive_drop_in or render or multi_device" 2>&1 | tail -8 > gpurun_out/r02o_pytest.txt
cat gpurun_out/r02o_pytest.txt
run() { tag=$1; wl=$2; res=$3; shift 3; env "$@" timeout 300 python bench.py --workload $wl --resolution $res --steps 3 --warmup 2 --no-cpu-baseline --no-extras > gpurun_out/r02o_$tag.json 2> gpurun_out/r02o_$tag.err; }
run c4_off c4 1024 BL_RAY_ORDER=0
run c4_on c4 1024 A=1
run c4_on_c7 c4 1024 BL_POL_OCC=3,7,5
run c4_on_c8 c4 1024 BL_POL_OCC=3,8,5
run sim_off simulation 1024 BL_RAY_ORDER=0
run sim_on simulation 1024 A=1
run formula_on formula 512 A=1
run tc_off true_color 512 BL_RAY_ORDER=0
run tc_on true_color 512 A=1
