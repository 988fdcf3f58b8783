#!/bin/bash
# compute-sanitizer pass over the GPU tests at small sizes (race / memory checking of the step kernels).
# usage (on a GPU box): bash tools/gpu_sanitize.sh [memcheck|initcheck|racecheck]
# Round 2: memcheck clean on k_step_fast2 (+ tie log), k_scatter_lean(_posuv), k_expand, k_scan_onepass, k_neigh_table_warp,
# k_hop_csr, k_seed, the host-buffer path (k_ingest32 / k_egest32_lean), the async export and the slab kernels; racecheck: 0 hazards
# on k_step_fast2; initcheck: clean on the single-context fast path (the slab transports copy fixed-size messages, whose tails are
# zeroed once at t2d_comm_init for that reason).  initcheck over the slab tests takes ~4 minutes.
TOOL=${1:-memcheck}
timeout 900 compute-sanitizer --tool $TOOL --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x \
    -k "fp32_fast_path or seeded or noise or coincident or empty or seeding or async_export or step_host" 2>&1 | tail -6
timeout 900 compute-sanitizer --tool $TOOL --error-exitcode 9 python -m pytest tests/test_gpu_fastpath.py -q -x -k "one_step_bar" 2>&1 | tail -4
timeout 900 compute-sanitizer --tool $TOOL --error-exitcode 9 python -m pytest tests/test_table_csr.py -q -x -k "csr_input or builder" 2>&1 | tail -4
timeout 600 compute-sanitizer --tool $TOOL --error-exitcode 9 python -m pytest tests/test_gpu_slabs.py -q -x 2>&1 | tail -6
