#!/bin/bash
# One bench.py line per BASELINE configuration (through the driver's door), written to gpurun_out/bench_r2_<cfg>.json
# (copy the ones to be judged into profiles/).  Usage: scripts/bench_all.sh [tag]   (run on the GPU box, from the repo root)
tag=${1:-r2}
mkdir -p gpurun_out
for cfg in C3 C1 C2 C4 C5; do
  steps=3; [ "$cfg" = C5 ] && steps=2
  python bench.py --config $cfg --steps $steps --warmup 3 > gpurun_out/bench_${tag}_${cfg}.json 2> gpurun_out/bench_${tag}_${cfg}.err \
    || echo "bench $cfg failed (see gpurun_out/bench_${tag}_${cfg}.err)"
  tail -c 400 gpurun_out/bench_${tag}_${cfg}.json; echo
done
