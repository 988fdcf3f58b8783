#!/bin/bash
# ncu --set full capture of selected kernels of the default bench. usage: gpu_ncu.sh TAG REGEX [SKIP] [COUNT]
TAG=${1:-n}; RE=${2:-k_step_euclid}; SKIP=${3:-5}; CNT=${4:-2}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$RE" -s $SKIP -c $CNT -o $OUT/prof python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1; echo "ncu rc=$?"; tail -3 $OUT/ncu_full.log
