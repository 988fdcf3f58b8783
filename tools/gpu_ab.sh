#!/bin/bash
# A/B of the step kernels on one box: fast-path GPU tests, then the default bench with k_step_fast2 and the legacy kernel,
# with and without the tie log.  usage: gpu_ab.sh TAG [pytest-args]
TAG=${1:-ab}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
(time timeout 900 python -m pytest ${@:-tests/test_gpu_fastpath.py tests/test_gpu_parity.py tests/test_gpu_slabs.py} -m gpu -q -x -s) > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|error|Error|assert|ties_cutoff" $OUT/pytest.log | tail -30
run() { name=$1; shift; env T2D_VERBOSE=1 "$@" timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > $OUT/bench_$name.json 2> $OUT/bench_$name.err; echo "$name: $(python -c "import json,sys; d=json.load(open('$OUT/bench_$name.json')); print(d.get('ms_per_step'), d.get('kernel_ms'), d.get('fault_mask'))" 2>&1 | tail -1)"; }
run fast2 T2D_COUNT_TIES=0
run fast2_ties T2D_COUNT_TIES=1
run legacy T2D_STEP=legacy T2D_COUNT_TIES=0
