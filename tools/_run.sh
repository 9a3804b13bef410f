timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1024.json 2> gpurun_out/bench_1024.err
timeout 300 python bench.py --workload polarized --resolution 512 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pol512.json 2> gpurun_out/bench_pol512.err
