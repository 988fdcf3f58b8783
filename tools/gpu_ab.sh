#!/bin/bash
# A/B of the step kernels on one box: GPU tests, then the default bench with the tiled and the legacy kernel.
TAG=${1:-ab}
OUT=gpurun_out/$TAG
mkdir -p $OUT
(time timeout 900 python -m pytest tests -m gpu -q -x -s) > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|error|Error|assert|ties_cutoff" $OUT/pytest.log | tail -30
for V in tiled legacy; do
  if [ $V = legacy ]; then export T2D_STEP=legacy; else unset T2D_STEP; fi
  T2D_VERBOSE=1 timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > $OUT/bench_$V.json 2> $OUT/bench_$V.err; echo "bench $V rc=$?"; cat $OUT/bench_$V.json; grep "t2d:" $OUT/bench_$V.err | head -3
done
unset T2D_STEP
