BL_GEO_BLOCKS=2 BLACKLIGHT_B200_LIB=$PWD/gpurun_tmp/lib_b192.so timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b192.json 2> gpurun_out/bench_b192.err
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1024.json 2> gpurun_out/bench_1024.err
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
