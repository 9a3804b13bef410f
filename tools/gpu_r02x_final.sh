#!/bin/bash
# Round 2, final single-GPU call: the whole GPU suite, smoke, the bench exactly as the driver runs it, the reference arm,
# and the ncu launch list of the bench command.
set -x
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 ) > gpurun_out/r02x_pytest.txt 2>&1
cat gpurun_out/r02x_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r02x_bench_driver.json 2> gpurun_out/r02x_bench_driver.err
tail -4 gpurun_out/r02x_bench_driver.err
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 ) > gpurun_out/r02x_bench_reference.json 2> gpurun_out/r02x_bench_reference.err
cut -c1-400 gpurun_out/r02x_bench_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02x_launches.csv python bench.py --steps 1 --warmup 0 --no-extras --no-cpu-baseline > gpurun_out/r02x_launches.log 2>&1
tail -2 gpurun_out/r02x_launches.log | cut -c1-200
