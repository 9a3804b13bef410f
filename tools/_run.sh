set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
for b in 2 3 4; do
  BL_GEO_BLOCKS=$b timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_geo$b.json 2> gpurun_out/bench_geo$b.err
done
timeout 300 python bench.py --workload polarized --resolution 512 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_pol512.json 2> gpurun_out/bench_pol512.err
