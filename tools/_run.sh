mkdir -p /tmp/p
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"radiate_unpol" -c 1 -f -o /tmp/p/u python bench.py --resolution 512 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_u.log 2>&1
ncu -i /tmp/p/u.ncu-rep --page raw --csv > gpurun_out/r01h_unpol_raw.csv
ncu -i /tmp/p/u.ncu-rep --page source --csv --print-source cuda,sass 2>/dev/null | gzip -9 > gpurun_out/r01h_unpol_source.csv.gz
