#!/usr/bin/env python
"""Determinism stress per pipeline prefix: several contexts (threads; ctypes drops the GIL) run
PREP..stage S in one call concurrently, then tap the result; every repetition must equal the first."""
import os
import sys
import threading
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import tea_stereo_matching_b200 as t
from tea_stereo_matching_b200 import _native as N
from tea_stereo_matching_b200.adcensus import StageRunner

z = np.load(Path(__file__).resolve().parents[1] / "tests/golden/pair_0600_320x180.npz")
l, r = z["left"], z["right"]
D = int(sys.argv[1]) if len(sys.argv) > 1 else 48
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
nthreads = int(sys.argv[3]) if len(sys.argv) > 3 else 3

PREFIX = [("init", N.STAGE_PREP | N.STAGE_INIT, "vol"),
          ("aggregate", N.STAGE_PREP | N.STAGE_INIT | N.STAGE_AGGREGATE, "vol"),
          ("scanline", N.STAGE_PREP | N.STAGE_INIT | N.STAGE_AGGREGATE | N.STAGE_SCANLINE, "vol+wta"),
          ("lrc", 63, "disp"), ("vote", 127, "disp"), ("interp", 255, "disp"), ("discont", 511, "disp"),
          ("all", N.STAGE_ALL, "final")]
if len(sys.argv) > 4:
    PREFIX = [p for p in PREFIX if p[0] in sys.argv[4].split(",")]


import ctypes as C
HASH = os.environ.get("HASH") == "1"


def hashes(run):
    out = []
    for w in range(100, 107):
        v = C.c_ulonglong(0)
        run.ctx._lib.tsm_selftest(run.ctx.handle, w, C.byref(v))
        out.append(v.value)
    return np.array(out, np.uint64)


def snapshot(run, kind):
    if HASH:
        return [hashes(run), run.volume(0)]
    if kind == "vol":
        return [run.volume(0), run.volume(1)]
    if kind == "vol+wta":
        return [run.volume(0), run.wta(0), run.wta(1)]
    if kind == "disp":
        return [run.disp()]
    return [run.final()]


import os
SPLIT = os.environ.get("SPLIT") == "1"
SCANONLY = len(sys.argv) > 4 and sys.argv[4] == "scanonly"
if SCANONLY:
    PREFIX = [("scanonly", N.STAGE_SCANLINE, "vol+wta")]
    agg_run = StageRunner(l, r, D)
    agg_run.run(N.STAGE_PREP | N.STAGE_INIT | N.STAGE_AGGREGATE)
    AGG = [agg_run.volume(0), agg_run.volume(1)]


def do_run(run, mask):
    if SCANONLY:
        run.run(N.STAGE_PREP)
        run.set_volume(0, AGG[0])
        run.set_volume(1, AGG[1])
    if SPLIT and mask & N.STAGE_SCANLINE:
        run.run(mask & ~N.STAGE_SCANLINE & ~(N.STAGE_ALL ^ 15))  # everything before the scanline, then a host sync
        run.run(mask & ~7)
    else:
        run.run(mask)


for name, mask, kind in PREFIX:
    base_run = StageRunner(l, r, D)
    do_run(base_run, mask)
    base = snapshot(base_run, kind)
    bad = [0]
    lock = threading.Lock()

    stop = threading.Event()
    NOISE = int(sys.argv[5]) if len(sys.argv) > 5 else 0  # mask run by all threads but thread 0, unchecked

    def worker(k):
        run = StageRunner(l, r, int(os.environ.get("NOISE_D", D)) if (NOISE and k > 0) else D)
        if NOISE and k > 0:
            run.run(N.STAGE_PREP | N.STAGE_INIT)
            while not stop.is_set():
                run.run(NOISE)
            run.ctx.synchronize() if hasattr(run.ctx, "synchronize") else None
            return
        for it in range(reps):
            do_run(run, mask)
            got = snapshot(run, kind)
            if not all(np.array_equal(x, y, equal_nan=True) for x, y in zip(got, base)):
                with lock:
                    bad[0] += 1
                    if bad[0] == 1 and kind == "vol+wta":
                        out = Path(__file__).resolve().parents[1] / "gpurun_out"
                        out.mkdir(exist_ok=True)
                        np.savez_compressed(out / "stress_dump.npz", got=got[0], base=base[0])
                    if bad[0] <= 2:
                        print(f"  thread {k} iter {it}: differs, counts {[int((x != y).sum()) for x, y in zip(got, base)]}", flush=True)
                        for x, y in zip(got, base):
                            w = np.argwhere(x != y)
                            if len(w):
                                print("    shape", x.shape, "min idx", w.min(0).tolist(), "max idx", w.max(0).tolist(),
                                      "first", w[:6].tolist(), "vals", x[tuple(w[0])], y[tuple(w[0])], flush=True)
                                if x.ndim == 3:
                                    ys, cnt = np.unique(w[:, 0], return_counts=True)
                                    print("    rows:", dict(zip(ys.tolist()[:12], cnt.tolist()[:12])), flush=True)

    ths = [threading.Thread(target=worker, args=(k,)) for k in range(nthreads)]
    [th.start() for th in ths]
    ths[0].join()
    stop.set()
    [th.join() for th in ths]
    print(f"prefix ..{name}: {bad[0]} of {reps * nthreads} differ", flush=True)
