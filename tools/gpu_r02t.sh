#!/bin/bash
# Round 2, twentieth GPU call: executed-flop counts of every kernel (profiles/executed_flops.json) and full ncu reports of the
# four polarized kernels (dense slab), the geodesic and the unpolarized kernel.
set -x
bash tools/ncu_capture.sh r02t flops
bash tools/ncu_capture.sh r02t full
