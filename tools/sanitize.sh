#!/bin/bash
# compute-sanitizer memcheck over small end-to-end cases of every kernel family (run under gpurun, one GPU):
# formula smoke, spherical / Cartesian / FMKS grids, inter-block interpolation, polarized, adaptive refinement.
# Writes gpurun_out/sanitizer_memcheck.txt; the last line of each block is the tool's ERROR SUMMARY.
out=gpurun_out/sanitizer_memcheck.txt
: > $out
run() {
  echo "== $*" >> $out
  timeout 500 compute-sanitizer --tool memcheck --print-limit 5 "$@" 2>&1 | grep -E "ERROR SUMMARY|Invalid|passed|failed|smoke ok|error" | tail -8 >> $out
}
run python -c "import __graft_entry__ as g; g.smoke()"
run python -m pytest tests -m gpu -q -x -k "test_iharm3d_reader_against_reference or test_athenak_reader_against_reference"
run python -m pytest tests -m gpu -q -x -k "test_live_reference_block_interpolation or test_adaptive_drop_in or test_golden_polarized or test_golden_render"
cat $out
