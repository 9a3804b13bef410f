#!/bin/bash
# Round 2, fourth GPU call (2 GPUs): new parity tests, polarized failure diagnosis, the bench at N = 1 and N = 2.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "polarized or multi_device" 2>&1 | tail -60 > gpurun_out/r02d_tests_pol.txt
tail -5 gpurun_out/r02d_tests_pol.txt
timeout 1500 python -m pytest tests -m gpu -q -k "cell_indices_exact or (full_resolution and c4)" -s 2>&1 | tail -40 > gpurun_out/r02d_tests_big.txt
tail -5 gpurun_out/r02d_tests_big.txt
CUDA_VISIBLE_DEVICES=0 timeout 900 python bench.py --resolution 2048 --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/r02d_bench_n1.json 2> gpurun_out/r02d_bench_n1.err
tail -3 gpurun_out/r02d_bench_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --resolution 2048 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02d_bench_n2.json 2> gpurun_out/r02d_bench_n2.err
tail -5 gpurun_out/r02d_bench_n2.err
# the drop-in executable on two devices against one device
python - <<'PY' > gpurun_out/r02d_driver.txt 2>&1
import sys, os, time, tempfile
import numpy as np
sys.path.insert(0, os.getcwd())
from blacklight_b200.cases import Case, C4_PHYSICS
d = tempfile.mkdtemp()
case = Case(d, 'simulation.input', dict(C4_PHYSICS, camera_resolution=1024))
t0 = time.time(); one, t1 = case.run_gpu_file(tag='one', devices=[0]); w1 = time.time() - t0
t0 = time.time(); two, t2 = case.run_gpu_file(tag='two', devices=[0, 1]); w2 = time.time() - t0
print('1 device: wall %.2f s %s' % (w1, t1))
print('2 devices: wall %.2f s %s' % (w2, t2))
print('bitwise equal:', all(np.array_equal(one[k], two[k], equal_nan=True) for k in one))
PY
cat gpurun_out/r02d_driver.txt
