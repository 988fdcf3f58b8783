#!/bin/bash
# sweep an environment knob over the default bench.  usage: gpu_sweep.sh TAG VAR v1 v2 ...
TAG=$1; VAR=$2; shift 2; OUT=gpurun_out/$TAG; mkdir -p $OUT
for v in "$@"; do
  env $VAR=$v python bench.py --steps 50 --warmup 5 --no-cpu-baseline > $OUT/bench_$v.json 2> $OUT/bench_$v.err; echo "$VAR=$v rc=$? $(cat $OUT/bench_$v.json)"
done
