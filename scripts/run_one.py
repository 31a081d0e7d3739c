#!/usr/bin/env python
"""Runs the ADCensus path a few times on one synthetic pair (for ncu / timing)."""
import argparse
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import tea_stereo_matching_b200 as t
from tea_stereo_matching_b200.synth import synth_v1

ap = argparse.ArgumentParser()
ap.add_argument("--H", type=int, default=1080)
ap.add_argument("--W", type=int, default=1920)
ap.add_argument("--D", type=int, default=192)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--gray", action="store_true")
ap.add_argument("--pair", default=None, help="npz with left/right (e.g. tests/golden/pair_c1_0600_720p.npz)")
a = ap.parse_args()
if a.pair:
    z = np.load(a.pair)
    l, r = z["left"], z["right"]
else:
    l, r = synth_v1(a.H, a.W, a.D, seed=1000, gray=a.gray)
m = t.ADCensus()
m.setMatchingStrategy(t.ColorModel.RGB)
m.setMinMaxDisparity(0, a.D)
m.context.set_profiling(True)
for i in range(a.reps):
    t0 = time.perf_counter()
    out = m.compute(l, r)
    dt = time.perf_counter() - t0
    st = m.context.stage_times()
    print(f"rep {i}: {dt*1e3:.2f} ms wall; stages(ms): " + ", ".join(f"{k}={v:.3f}" for k, v in st.items()), flush=True)
print("valid fraction", float((out >= 0).mean()))
