#!/bin/bash
# Round 2, twenty-second GPU call: transfer kernel with several frequencies per thread (A/B).
set -x
mkdir -p gpurun_out
run() { tag=$1; wl=$2; res=$3; shift 3; env "$@" timeout 300 python bench.py --workload $wl --resolution $res --steps 3 --warmup 2 --no-cpu-baseline --no-extras > gpurun_out/r02v_$tag.json 2> gpurun_out/r02v_$tag.err; }
run c4_base c4 1024 A=1
run c4_m4_3 c4 1024 BL_POL_MULTI=4,3
run c4_m4_2 c4 1024 BL_POL_MULTI=4,2
run c4_m4_4 c4 1024 BL_POL_MULTI=4,4
run c4_m2_3 c4 1024 BL_POL_MULTI=2,3
run c4_m2_4 c4 1024 BL_POL_MULTI=2,4
run c4_m2_5 c4 1024 BL_POL_MULTI=2,5
BL_POL_MULTI=4,3 timeout 600 python -m pytest tests -m gpu -q -x -k "golden_polarized or pipeline_matches" 2>&1 | tail -4
