#!/usr/bin/env python
"""Determinism stress: the same pair through several contexts concurrently, many times; every map must be identical."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import tea_stereo_matching_b200 as t

z = np.load(Path(__file__).resolve().parents[1] / "tests/golden/pair_0600_320x180.npz")
l, r = z["left"], z["right"]
D = int(sys.argv[1]) if len(sys.argv) > 1 else 48
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
ms = []
for k in range(3):
    m = t.ADCensus()
    m.setMatchingStrategy(t.ColorModel.RGB)
    m.setMinMaxDisparity(0, D)
    ms.append(m)
ref = ms[0].compute(l, r)
bad = 0
for it in range(reps):
    for m in ms:
        m.enqueue(l, r)
    outs = [m.wait() for m in ms]
    for k, o in enumerate(outs):
        if not np.array_equal(o, ref):
            bad += 1
            d = np.argwhere(o != ref)
            print(f"iter {it} ctx {k}: {len(d)} pixels differ, first {d[:5].tolist()}", flush=True)
print("mismatching maps:", bad, "of", reps * len(ms))
sys.exit(1 if bad else 0)
