set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1024.json 2> gpurun_out/bench_1024.err
for b in 2 4; do
  BL_GEO_BLOCKS=$b timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_geo$b.json 2> gpurun_out/bench_geo$b.err
done
