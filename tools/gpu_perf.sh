#!/bin/bash
# perf iteration: default bench without the CPU/e2e legs (+ optional ncu source capture of the step kernel)
# usage: gpu_perf.sh TAG [ncu]
TAG=${1:-p}; OUT=gpurun_out/$TAG; mkdir -p $OUT
python bench.py --steps 50 --warmup 5 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json; tail -3 $OUT/bench.err
if [ "$2" = "ncu" ]; then bash tools/gpu_ncu.sh $TAG k_step_euclid 50 1 > $OUT/ncu.log 2>&1; tail -2 $OUT/ncu.log; fi
