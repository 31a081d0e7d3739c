#!/bin/bash
# A/B of the aggregation ring-home mix inside one gpurun call
for rep in 1 2; do
for mix in 0 "4,4,4,4" "4,6,4,6" "3,6,4,6" "4,5,4,5" "4,6,4,4" "4,4,4,6" "2,6,2,6"; do
  echo -n "mix $mix: "
  TSM_AGG_MIX=$mix timeout 120 python scripts/run_one.py --D ${ABD:-192} --reps 4 | tail -2 | head -1 | sed 's/.*aggregate=\([0-9.]*\).*/aggregate \1 ms/'
done
done
