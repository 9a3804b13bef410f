#!/bin/bash
# Round-end check on a B200 (run under gpurun, one GPU): the parity suite, the contract bench with its CPU baseline,
# the reference arm, the ncu launch list of the bench command (per-launch times under ncu are cold-cache and
# serialised: only the kernels' shares of the step are comparable with bench.py's numbers) and the smoke test.
set -x
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 400 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
