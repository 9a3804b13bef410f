#!/bin/bash
# Round 2, twenty-first GPU call: transfer kernel with the transport matrix staged through shared memory (A/B, parity).
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -k "polarized or cks or ray_ordering or full_resolution" 2>&1 | tail -8 > gpurun_out/r02u_pytest.txt
cat gpurun_out/r02u_pytest.txt
run() { tag=$1; wl=$2; res=$3; shift 3; env "$@" timeout 300 python bench.py --workload $wl --resolution $res --steps 3 --warmup 2 --no-cpu-baseline --no-extras > gpurun_out/r02u_$tag.json 2> gpurun_out/r02u_$tag.err; }
run c4_tile0 c4 1024 BL_POL_TILE=0
run c4_tile1 c4 1024 A=1
run c4_tile1_t6 c4 1024 BL_POL_OCC=3,0,6,5
run c4_tile1_t4 c4 1024 BL_POL_OCC=3,0,4,5
run polth_tile1 polarized_thermal 1024 A=1
run c4_2048 c4 2048 A=1
