#!/bin/bash
# Round 2, eleventh GPU call: deferred emission of the DP integrator -- bit parity, then A/B.
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -k "golden_unpolarized or live_reference_unpolarized or checkpoint or waves or golden_polarized or division or adaptive_drop_in" 2>&1 | tail -12
run() { tag=$1; wl=$2; res=$3; shift 3; env "$@" timeout 300 python bench.py --workload $wl --resolution $res --steps 3 --warmup 2 --no-cpu-baseline --no-extras > gpurun_out/r02k_$tag.json 2> gpurun_out/r02k_$tag.err; }
run formula_off formula 512 BL_GEO_DEFER=0
run formula_on formula 512
run formula_on_min2 formula 512 BL_GEO_DEFER_MIN=2
run formula_on_min5 formula 512 BL_GEO_DEFER_MIN=5
run formula_on_g2 formula 512 BL_GEO_BLOCKS=2
run formula_1024_on formula 1024
run sim_off simulation 1024 BL_GEO_DEFER=0
run sim_on simulation 1024
run sim_on_min2 simulation 1024 BL_GEO_DEFER_MIN=2
run sim_on_min5 simulation 1024 BL_GEO_DEFER_MIN=5
run sim_on_g2 simulation 1024 BL_GEO_BLOCKS=2
run c4_on c4 1024
run c4_c5 c4 1024 BL_POL_OCC=3,5,5
