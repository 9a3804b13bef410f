#!/bin/bash
# Round 2, ninth GPU call: record prefetch and geodesic barrier A/B.
set -x
mkdir -p gpurun_out
run() { tag=$1; wl=$2; res=$3; shift 3; env "$@" timeout 300 python bench.py --workload $wl --resolution $res --steps 3 --warmup 2 --no-cpu-baseline --no-extras > gpurun_out/r02i_$tag.json 2> gpurun_out/r02i_$tag.err; }
run sim_base simulation 1024 BL_RAD_PREFETCH=0
run sim_pf1 simulation 1024 BL_RAD_PREFETCH=1
run sim_pf2 simulation 1024 BL_RAD_PREFETCH=2
run sim_pf4 simulation 1024 BL_RAD_PREFETCH=4
run sim_nosync simulation 1024 BL_GEO_SYNC=0
run c4_base c4 1024 BL_RAD_PREFETCH=0
run c4_pf2 c4 1024 BL_RAD_PREFETCH=2
run formula_base formula 512 BL_GEO_SYNC=1
run formula_nosync formula 512 BL_GEO_SYNC=0
run formula_nosync_g2 formula 512 BL_GEO_SYNC=0 BL_GEO_BLOCKS=2
run formula_1024_base formula 1024 BL_GEO_SYNC=1
run formula_1024_nosync formula 1024 BL_GEO_SYNC=0
timeout 300 python -m pytest tests -m gpu -q -x -k "golden_unpolarized or golden_polarized" 2>&1 | tail -3
BL_GEO_SYNC=0 BL_RAD_PREFETCH=2 timeout 300 python -m pytest tests -m gpu -q -x -k "golden_unpolarized or golden_polarized" 2>&1 | tail -3
