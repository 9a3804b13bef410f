#!/bin/bash
# Round 2, fifth GPU call (1 GPU): repaired and new parity tests, variants of the geometry / transfer stages.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s -k "pipeline_matches or multi_device or roundoff_limited" 2>&1 | grep -v "^$" | tail -40 > gpurun_out/r02e_tests_a.txt
tail -4 gpurun_out/r02e_tests_a.txt
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --resolution 1024 --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r02e_$tag.json 2> gpurun_out/r02e_$tag.err; }
run base BL_POL_OCC=3,4,5,0,0
run gsync BL_POL_OCC=3,4,5,1,0
run g2sync BL_POL_OCC=2,4,5,1,0
run tpf5 BL_POL_OCC=3,4,5,0,1
run tpf4 BL_POL_OCC=3,4,4,0,1
timeout 1700 python -m pytest tests -m gpu -q -s -k "cell_indices_exact or (full_resolution and c4)" 2>&1 | grep -v "^$" | tail -60 > gpurun_out/r02e_tests_b.txt
tail -6 gpurun_out/r02e_tests_b.txt
