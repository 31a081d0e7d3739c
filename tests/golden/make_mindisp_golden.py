#!/usr/bin/env python
"""Golden vectors for setMinMaxDisparity(min > 0, max): every stage output of the UNMODIFIED reference
(oracle/_ref/libadcensus_ref.so, serial-scanline semantics) on a 48 x 128 crop of the demo pair with (min, max) = (4, 28),
full volumes included, so the CUDA path can be checked stage by stage on a box without /root/reference.

    python tests/golden/make_mindisp_golden.py      ->  tests/golden/ref_0600_crop_128x48_d4_28.npz
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import oracle  # noqa: E402

OUT = Path(__file__).resolve().parent
MIND, MAXD = 4, 28


def main():
    oracle.build()
    ref = oracle.Ref()
    z = np.load(OUT / "pair_0600_320x180.npz")
    left = np.ascontiguousarray(z["left"][60:108, 100:228])
    right = np.ascontiguousarray(z["right"][60:108, 100:228])
    st = ref.run(left, right, MAXD, serial_scanline=True, min_disp=MIND)
    d = dict(left=left, right=right, min_disparity=MIND, max_disparity=MAXD)
    for name in ("vol_init", "vol_agg", "vol_scan"):
        for k in range(2):
            d[f"{name}{k}"] = getattr(st, name)[k]
    for k in range(2):
        d[f"arms{k}"] = np.stack(st.arms[k], axis=2).astype(np.uint8)
        d[f"wta{k}"] = st.wta[k].astype(np.int16)
    d["lrc"] = st.lrc.astype(np.int16)
    for i in range(5):
        d[f"vote{i}"] = st.vote[i].astype(np.int16)
    d["interp"] = st.interp.astype(np.int16)
    d["discont"] = st.discont.astype(np.int16)
    d["final"] = st.final
    np.savez_compressed(OUT / f"ref_0600_crop_128x48_d{MIND}_{MAXD}.npz", **d)
    print("valid fraction", float((st.final >= MIND).mean()), "wta range", int(st.wta[0].min()), int(st.wta[0].max()))


if __name__ == "__main__":
    main()
