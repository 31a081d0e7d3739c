#!/usr/bin/env python
"""Times the HSI colour model (the reference's default-constructed state) on one synthetic 1080p pair."""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import tea_stereo_matching_b200 as t
from tea_stereo_matching_b200.synth import synth_v1

l, r = synth_v1(1080, 1920, 192, seed=1000)
m = t.ADCensus()  # default: HSI
m.setMinMaxDisparity(0, 192)
t0 = time.perf_counter()
m.compute(l, r)
print("first HSI call (2^24-entry bgr2hsi table built on the host + arena): %.2f s" % (time.perf_counter() - t0))
m.context.set_profiling(True)
for i in range(3):
    t0 = time.perf_counter()
    out = m.compute(l, r)
    dt = time.perf_counter() - t0
print("HSI 1080p D=192: %.2f ms" % (dt * 1e3), {k: round(v, 3) for k, v in m.context.stage_times().items()})
print("valid fraction", float((out >= 0).mean()))
