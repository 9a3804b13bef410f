set -x
timeout 900 python -m pytest tests -m gpu -q -k "polar or adaptive or slow or code_kappa or harm3d or block_interp or cartesian" 2>&1 | tail -4
timeout 300 python bench.py --workload polarized_thermal --resolution 1024 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_polth.json 2> gpurun_out/bench_polth.err
timeout 300 python bench.py --workload polarized --resolution 512 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pol512.json 2> gpurun_out/bench_pol512.err
