#!/bin/bash
# Round 2, sixth GPU call (1 GPU): the whole parity suite after the table-driven exp / log, then the bench.
set -x
mkdir -p gpurun_out
timeout 300 python bench.py --resolution 1024 --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r02f_c4_1024.json 2> gpurun_out/r02f_c4_1024.err
timeout 300 python bench.py --workload simulation --steps 3 --warmup 2 --no-cpu-baseline --no-extras > gpurun_out/r02f_sim_1024.json 2> gpurun_out/r02f_sim_1024.err
timeout 300 python bench.py --workload polarized_thermal --steps 3 --warmup 2 --no-cpu-baseline --no-extras > gpurun_out/r02f_polth_1024.json 2> gpurun_out/r02f_polth_1024.err
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | grep -v "^$" | tail -60 > gpurun_out/r02f_tests.txt
tail -8 gpurun_out/r02f_tests.txt
