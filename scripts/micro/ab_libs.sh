#!/bin/bash
# A/B of library variants inside one gpurun call: every lib in scripts/micro/libs against the default build.
# Prints the per-stage and per-launch times (ms) of the last of 4 repetitions.
fmt() { tail -2 | head -1 | sed 's/.*stages(ms): //' | tr ',' '\n' | grep -E "cost_init|aggregate|scanline" | grep -v "#n" | tr -d ' ' | tr '\n' ' '; echo; }
for rep in 1 2; do
for lib in scripts/micro/libs/*.so; do
  echo -n "$(basename $lib .so): "
  TSM_LIB=$PWD/$lib timeout 120 python scripts/run_one.py --D ${ABD:-192} --reps 4 | fmt
done
echo -n "default: "; timeout 120 python scripts/run_one.py --D ${ABD:-192} --reps 4 | fmt
done
