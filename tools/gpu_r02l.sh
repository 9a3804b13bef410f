#!/bin/bash
# Round 2, twelfth GPU call: full GPU suite after removing deferred emission, the default bench line (4096^2, N = 1), formula occupancy.
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r02l_pytest.txt
cat gpurun_out/r02l_pytest.txt
( time timeout 1200 python bench.py --steps 3 --warmup 3 ) > gpurun_out/r02l_bench_default.json 2> gpurun_out/r02l_bench_default.err
tail -4 gpurun_out/r02l_bench_default.err
run() { tag=$1; wl=$2; res=$3; shift 3; env "$@" timeout 300 python bench.py --workload $wl --resolution $res --steps 3 --warmup 2 --no-cpu-baseline --no-extras > gpurun_out/r02l_$tag.json 2> gpurun_out/r02l_$tag.err; }
run formula_1024_g2 formula 1024 BL_GEO_BLOCKS=2
run sim_g2 simulation 1024 BL_GEO_BLOCKS=2
