#!/bin/bash
# Builds library variants for A/B runs on the GPU box: scripts/micro/libs/<name>.so = the default objects with ONE source
# recompiled with extra macros.  Usage: scripts/micro/build_variants.sh <source.cu> name1 "-DX=1 -DY=2" [name2 "..."] ...
set -e
cd "$(dirname "$0")/../../tea_stereo_matching_b200/csrc"
make -s > /dev/null
src=$1; shift
base=${src%.cu}
mkdir -p ../../scripts/micro/libs build/var
ARCH="-gencode arch=compute_100a,code=sm_100a"
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  nvcc -O3 -std=c++17 $ARCH -lineinfo -fmad=false -Xcompiler -fPIC -Xptxas -v $flags -c $src -o build/var/${base}_$name.o 2> build/var/${base}_$name.log
  objs=""
  for o in tsm_capi k_prep k_cost k_aggregate k_scanline k_scanline3 k_post k_rectify k_consumers; do
    if [ $o = $base ]; then objs="$objs build/var/${base}_$name.o"; else objs="$objs build/$o.o"; fi
  done
  nvcc $ARCH -shared -o ../../scripts/micro/libs/$name.so $objs -lpthread
  echo "built scripts/micro/libs/$name.so ($flags): $(grep -A2 'k_agg_fused\|k_scanline\|k_cost_init' build/var/${base}_$name.log | grep -o 'Used [0-9]* registers' | sort -u | tr '\n' ' ')"
done
