#!/bin/bash
TAG=${1:-pt}
OUT=gpurun_out/$TAG
mkdir -p $OUT
run() { name=$1; shift; env T2D_VERBOSE=1 "$@" timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > $OUT/bench_$name.json 2> $OUT/bench_$name.err; echo "$name: $(cat $OUT/bench_$name.json) $(grep 't2d:' $OUT/bench_$name.err | head -1)"; }
run legacy_noties T2D_STEP=legacy T2D_COUNT_TIES=0
run tiled_noties T2D_COUNT_TIES=0
run tiled T2D_COUNT_TIES=1
run tiled_tc32 T2D_TILE_CELLS=32 T2D_COUNT_TIES=0
run tiled_tc128 T2D_TILE_CELLS=128 T2D_COUNT_TIES=0
run tiled_gap0 T2D_TILE_GAP=0 T2D_COUNT_TIES=0
run tiled_gap32 T2D_TILE_GAP=32 T2D_COUNT_TIES=0
timeout 600 python -m pytest tests/test_gpu_fastpath.py -x -q > $OUT/pytest.log 2>&1; tail -3 $OUT/pytest.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step_euclid_tiled -s 40 -c 1 -f -o $OUT/prof python bench.py --steps 2 --warmup 40 --no-cpu-baseline > $OUT/ncu.log 2>&1; echo "ncu rc=$?"
