#!/usr/bin/env python
"""Determinism stress with DIFFERENT pairs in flight (computeBatch, four contexts per GPU): every map of every repetition must
equal the map of the same pair computed alone.  Usage: stress_mixed.py [D] [reps]"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import tea_stereo_matching_b200 as t
from tea_stereo_matching_b200.synth import synth_v1

z = np.load(Path(__file__).resolve().parents[1] / "tests/golden/pair_0600_320x180.npz")
D = int(sys.argv[1]) if len(sys.argv) > 1 else 48
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
pairs = [(z["left"], z["right"])] + [synth_v1(180, 320, D, seed=70 + i) for i in range(6)]
m = t.ADCensus(device=0)
m.setMatchingStrategy(t.ColorModel.RGB)
m.setMinMaxDisparity(0, D)
single = [m.compute(l, r) for l, r in pairs]
bad = 0
for it in range(reps):
    outs = m.computeBatch([p[0] for p in pairs], [p[1] for p in pairs])
    for k, (a, b) in enumerate(zip(single, outs)):
        if not np.array_equal(a, b):
            bad += 1
            d = np.argwhere(a != b)
            print(f"iter {it} pair {k}: {len(d)} pixels differ, first {d[:4].tolist()}", flush=True)
print("mismatching maps:", bad, "of", reps * len(pairs))
sys.exit(1 if bad else 0)
