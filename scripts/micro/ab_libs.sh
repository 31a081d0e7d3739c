#!/bin/bash
# A/B of library variants inside one gpurun call: every lib in scripts/micro/libs against the default build
for rep in 1 2; do
for lib in scripts/micro/libs/*.so; do
  echo -n "$lib: "
  TSM_LIB=$PWD/$lib timeout 120 python scripts/run_one.py --D ${ABD:-192} --reps 4 | tail -2 | head -1 | sed 's/.*cost_init=\([0-9.]*\), aggregate=\([0-9.]*\), scanline=\([0-9.]*\).*/cost \1 aggregate \2 scanline \3/'
done
echo -n "default: "; timeout 120 python scripts/run_one.py --D ${ABD:-192} --reps 4 | tail -2 | head -1 | sed 's/.*cost_init=\([0-9.]*\), aggregate=\([0-9.]*\), scanline=\([0-9.]*\).*/cost \1 aggregate \2 scanline \3/'
done
