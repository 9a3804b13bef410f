#!/bin/bash
# Round 2, fourteenth GPU call: term-major kappa coefficients (shared exponentials across frequencies), table-driven logs per sample.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -k "polarized or full_resolution or cks or live_reference" 2>&1 | tail -8 > gpurun_out/r02n_pytest.txt
cat gpurun_out/r02n_pytest.txt
run() { tag=$1; wl=$2; res=$3; shift 3; env "$@" timeout 300 python bench.py --workload $wl --resolution $res --steps 3 --warmup 2 --no-cpu-baseline --no-extras > gpurun_out/r02n_$tag.json 2> gpurun_out/r02n_$tag.err; }
run c4_c4 c4 1024 BL_POL_OCC=3,4,5
run c4_c5 c4 1024 BL_POL_OCC=3,5,5
run c4_c6 c4 1024 BL_POL_OCC=3,6,5
run polth polarized_thermal 1024 A=1
run adaptive adaptive 512 A=1
