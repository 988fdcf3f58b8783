#!/bin/bash
# build a variant of the library with extra nvcc flags into variants/lib_<name>.so (dev tool for A/B runs)
# usage: build_variant.sh NAME "-DFOO=1 ..."
set -e
NAME=$1; FLAGS=$2
ROOT=$(cd $(dirname $0)/.. && pwd)
mkdir -p $ROOT/variants
make -s -C $ROOT/2dtissue_b200/csrc OBJDIR=$ROOT/variants/build_$NAME OUT=$ROOT/variants/lib_$NAME.so EXTRA="$FLAGS" -j8
grep -A3 "k_step_fast2ILb1ELb0" $ROOT/variants/build_$NAME/step_f32.ptxas.log | grep "Used\|spill"
