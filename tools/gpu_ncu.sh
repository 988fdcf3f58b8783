#!/bin/bash
# ncu --set full capture of selected kernels of the default bench + source/raw pages as CSV.  The workload densifies
# over the first tens of steps, so the capture is taken after WARM steps (default 50), like the bench's timed region.
# usage: gpu_ncu.sh TAG REGEX [WARM] [COUNT] [extra bench args...]
TAG=${1:-n}; RE=${2:-k_step_euclid}; WARM=${3:-50}; CNT=${4:-1}; shift 4
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$RE" -s $WARM -c $CNT -f -o $OUT/prof python bench.py --steps 2 --warmup $WARM --no-cpu-baseline "$@" > $OUT/ncu_full.log 2>&1; echo "ncu rc=$?"; tail -3 $OUT/ncu_full.log
ncu -i $OUT/prof.ncu-rep --page raw --csv > $OUT/raw.csv 2>/dev/null
ncu -i $OUT/prof.ncu-rep --page source --csv --print-source cuda,sass > $OUT/source.csv 2>$OUT/source.err || ncu -i $OUT/prof.ncu-rep --page source --csv > $OUT/source.csv 2>>$OUT/source.err
ls -la $OUT
