#!/bin/bash
# dev: ablation timing of k_step_fast2 (variants/lib_abl.so built with -DT2D_F2_ABLATE): the profile steps at the end of the
# bench run with parts of the kernel switched off.  usage: gpu_ablate.sh TAG
TAG=${1:-abl}; OUT=gpurun_out/$TAG; mkdir -p $OUT
cp 2dtissue_b200/lib2dtissue_b200.so /tmp/lib_main.so; cp variants/lib_abl.so 2dtissue_b200/lib2dtissue_b200.so
for m in ${MODES:-0 1 2 3}; do
  T2D_COUNT_TIES=0 T2D_F2_ABLATE_AFTER=55 T2D_F2_ABLATE_MODE=$m python bench.py --steps 50 --warmup 5 --no-cpu-baseline > $OUT/bench_$m.json 2> $OUT/bench_$m.err; echo "mode $m: $(cut -c1-200 $OUT/bench_$m.json)"
done
cp /tmp/lib_main.so 2dtissue_b200/lib2dtissue_b200.so
