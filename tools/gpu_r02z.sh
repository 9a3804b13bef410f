#!/bin/bash
# exp_bf scaling tweak: A/B timing + parity subset
set -x
mkdir -p gpurun_out
run() { tag=$1; wl=$2; res=$3; shift 3; env "$@" timeout 300 python bench.py --workload $wl --resolution $res --steps 3 --warmup 2 --no-cpu-baseline --no-extras > gpurun_out/r02z_$tag.json 2> gpurun_out/r02z_$tag.err; }
run c4 c4 1024 A=1
run sim simulation 1024 A=1
run polth polarized_thermal 1024 A=1
timeout 600 python -m pytest tests -m gpu -q -x -k "golden_polarized or golden_unpolarized or pipeline_matches or cks" 2>&1 | tail -3
