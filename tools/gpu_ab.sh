#!/bin/bash
# A/B two builds of the library with the default bench (dev tool). usage: gpu_ab.sh TAG path_to_variant_B.so
TAG=$1; B=$2; OUT=gpurun_out/$TAG; mkdir -p $OUT
python bench.py --steps 50 --warmup 5 --no-cpu-baseline > $OUT/A.json 2>$OUT/A.err; cat $OUT/A.json
cp 2dtissue_b200/lib2dtissue_b200.so /tmp/libA.so; cp $B 2dtissue_b200/lib2dtissue_b200.so
python bench.py --steps 50 --warmup 5 --no-cpu-baseline > $OUT/B.json 2>$OUT/B.err; cat $OUT/B.json
cp /tmp/libA.so 2dtissue_b200/lib2dtissue_b200.so
