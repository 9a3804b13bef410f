#!/bin/bash
# Round 2, thirteenth GPU call: device-side camera (bitwise vs host), drop-in with device camera, e2e vs resident.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -k "device_camera or drop_in or multi_device or adaptive or smoke" 2>&1 | tail -8 > gpurun_out/r02m_pytest.txt
cat gpurun_out/r02m_pytest.txt
run() { tag=$1; wl=$2; res=$3; shift 3; env "$@" timeout 300 python bench.py --workload $wl --resolution $res --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02m_$tag.json 2> gpurun_out/r02m_$tag.err; }
run sim_dev simulation 1024 A=1
run c4_dev c4 1024 A=1
timeout 300 python bench.py --workload simulation --resolution 1024 --steps 5 --warmup 3 --no-cpu-baseline --no-extras --host-camera > gpurun_out/r02m_sim_host.json 2> gpurun_out/r02m_sim_host.err
run adaptive_g3 adaptive 512 A=1
run adaptive_g2 adaptive 512 BL_GEO_BLOCKS=2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
