for v in p3 p4; do
  BLACKLIGHT_B200_LIB=$PWD/gpurun_tmp/lib_$v.so timeout 300 python bench.py --workload polarized --resolution 512 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err
done
