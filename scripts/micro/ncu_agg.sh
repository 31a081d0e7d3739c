#!/bin/bash
# per-kernel durations of the aggregation passes for a few configurations (ncu launch list; cold-cache times)
for D in 191 192; do for mix in 0 "4,4,4,6"; do
  TSM_AGG_MIX=$mix ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_agg --csv --log-file gpurun_out/agg_${D}_${mix//,/}.csv python scripts/run_one.py --D $D --reps 1 > /dev/null 2>&1
  echo "D=$D mix=$mix"; python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/agg_${D}_${mix//,/}.csv")) if len(r)>5 and r[0].isdigit()]
print(" ".join(f"{float(r[-1])/1e6:.3f}" for r in rows))
PY
done; done
