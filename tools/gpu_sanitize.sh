#!/bin/bash
# compute-sanitizer pass over the GPU tests at small sizes (race / memory checking of the step kernels).
# usage (on a GPU box): bash tools/gpu_sanitize.sh [memcheck|initcheck|racecheck]
# Round 1: memcheck and initcheck clean on the selections below (k_step_euclid_fast/exact, scan, scatter, slab kernels).
TOOL=${1:-memcheck}
timeout 900 compute-sanitizer --tool $TOOL --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x \
    -k "fp32_fast_path or seeded or noise or coincident or empty" 2>&1 | tail -6
timeout 600 compute-sanitizer --tool $TOOL --error-exitcode 9 python -m pytest tests/test_gpu_slabs.py -q -x 2>&1 | tail -6
