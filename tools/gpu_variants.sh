#!/bin/bash
# bench every variants/lib_*.so (dev tool). usage: gpu_variants.sh TAG [names...]
TAG=$1; shift; OUT=gpurun_out/$TAG; mkdir -p $OUT
cp 2dtissue_b200/lib2dtissue_b200.so /tmp/lib_main.so
NAMES="$@"; [ -z "$NAMES" ] && NAMES=$(ls variants/lib_*.so | sed 's/.*lib_\(.*\)\.so/\1/')
for n in main $NAMES; do
  if [ $n = main ]; then cp /tmp/lib_main.so 2dtissue_b200/lib2dtissue_b200.so; else cp variants/lib_$n.so 2dtissue_b200/lib2dtissue_b200.so; fi
  T2D_COUNT_TIES=0 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > $OUT/bench_$n.json 2> $OUT/bench_$n.err; echo "$n rc=$? $(cut -c1-200 $OUT/bench_$n.json)"
done
cp /tmp/lib_main.so 2dtissue_b200/lib2dtissue_b200.so
