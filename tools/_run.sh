set -x
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
timeout 600 python bench.py --workload polarized --resolution 512 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pol512.json 2> gpurun_out/bench_pol512.err
timeout 600 python bench.py --workload formula --resolution 512 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_formula512.json 2> gpurun_out/bench_formula512.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r01_launches_bench_1024.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
