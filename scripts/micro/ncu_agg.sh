#!/bin/bash
# per-kernel durations of the aggregation passes (ncu launch list; cold-cache, serialised times)
D=${1:-192}
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_agg --csv --log-file gpurun_out/agg_list.csv python scripts/run_one.py --D $D --reps 1 > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/agg_list.csv")) if len(r)>5 and r[0].isdigit()]
import collections
acc=collections.OrderedDict()
for r in rows:
    acc.setdefault((r[0], r[4][:40]), {})[r[-3]] = r[-1]
for (i,k),v in acc.items():
    print(i, k, {m.split('.')[0][-14:]: x for m,x in v.items()})
PY
