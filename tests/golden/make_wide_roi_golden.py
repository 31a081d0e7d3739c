#!/usr/bin/env python
"""Goldens for ROI / mask matching at the reference's own demo WIDTH (1280 px -> W / 2 + 1 = 641 disparity levels,
ADCensus.cpp:339-340): final maps of the UNMODIFIED reference's public compute() (one OpenMP thread = deterministic
scanline) on a 1280 x 96 stripe of demo-imgs/0600 (rows 300..395; the full 720 rows would take the single-threaded
reference about half an hour per mode).

    python tests/golden/make_wide_roi_golden.py   ->  tests/golden/ref_0600_stripe_1280x96_roi.npz   (needs /root/reference)
"""
import sys
from pathlib import Path

import cv2
import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import oracle  # noqa: E402

OUT = Path(__file__).resolve().parent
REF_IMGS = Path("/root/reference/demo-imgs")


def main():
    oracle.build()
    ref = oracle.Ref()
    L = cv2.imread(str(REF_IMGS / "0600-Left.bmp"))[300:396].copy()
    R = cv2.imread(str(REF_IMGS / "0600-Right.bmp"))[300:396].copy()
    # mask matching: black pixels are holes in both images (a masked foreground, as the mode is meant for)
    Lm, Rm = L.copy(), R.copy()
    Lm[:, :150] = 0; Rm[:, :110] = 0
    Lm[30:70, 600:700] = 0; Rm[30:70, 560:660] = 0
    np.savez_compressed(OUT / "ref_0600_stripe_1280x96_roi.npz", left=L, right=R, left_mask=Lm, right_mask=Rm,
                        rgb_roi_off5=ref.compute_ex(L, R, 64, "RGB", roi=True, offset=5),
                        hsi_mask_off0=ref.compute_ex(Lm, Rm, 64, "HSI", mask=True, offset=0))


if __name__ == "__main__":
    main()
