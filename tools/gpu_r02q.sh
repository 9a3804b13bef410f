#!/bin/bash
# Round 2, seventeenth GPU call: transfer-stage prefetch, integrator queue ordered by impact parameter.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -k "golden_unpolarized or live_reference_unpolarized or checkpoint or waves or golden_polarized or adaptive_drop_in or device_camera or ray_ordering" 2>&1 | tail -8 > gpurun_out/r02q_pytest.txt
cat gpurun_out/r02q_pytest.txt
run() { tag=$1; wl=$2; res=$3; shift 3; env "$@" timeout 300 python bench.py --workload $wl --resolution $res --steps 3 --warmup 2 --no-cpu-baseline --no-extras > gpurun_out/r02q_$tag.json 2> gpurun_out/r02q_$tag.err; }
run c4_pf0 c4 1024 BL_POL_PREFETCH=0
run c4_pf1 c4 1024 BL_POL_PREFETCH=1
run c4_pf2 c4 1024 BL_POL_PREFETCH=2
run c4_go0 c4 1024 BL_GEO_ORDER=0
run formula_go0 formula 512 BL_GEO_ORDER=0
run formula_go1 formula 512 A=1
run formula_go1_g2 formula 512 BL_GEO_BLOCKS=2
run formula1024_go1 formula 1024 A=1
run sim_go1 simulation 1024 A=1
