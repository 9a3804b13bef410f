#!/bin/bash
# compute-sanitizer over the kernels added in round 2 (run under gpurun, one GPU): camera pixels on the device, ray lists
# (shared-memory histograms / cursors, warp-level ranks), the four-kernel polarized pipeline addressed by list position,
# the integrator's ordered queue, sharded adaptive assembly.  memcheck over small end-to-end cases, racecheck over the
# sorting and pipeline kernels.  Writes gpurun_out/r02_sanitizer_{memcheck,racecheck}.txt.
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  out=gpurun_out/r02_sanitizer_$tool.txt
  : > $out
  run() {
    echo "== $*" >> $out
    timeout 900 compute-sanitizer --tool $tool --print-limit 5 "$@" 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Invalid|hazard|passed|failed|smoke ok|error" | tail -8 >> $out
  }
  run python -c "import __graft_entry__ as g; g.smoke()"
  run python -m pytest tests -m gpu -q -x -k "test_ray_ordering_does_not_change_a_bit"
  if [ $tool = memcheck ]; then
    run python -m pytest tests -m gpu -q -x -k "test_device_camera_is_bitwise_the_host_camera or test_drop_in_with_device_camera or test_adaptive_sharded_over_ranks or test_golden_polarized"
  fi
  cat $out
done
