#!/usr/bin/env python
"""Makes tests/golden/rectmap_cv_golden.npz: cv2.initUndistortRectifyMap(..., CV_16SC2) outputs for a few
calibrations (what stereo::EpipolarRectifyMap::compute calls, source/stereo_utils.cpp:157-169), plus a stereo
calibration YAML in the reference's format (tests/golden/stereo_calib.yml) written with cv2.FileStorage."""
from pathlib import Path

import cv2
import numpy as np

here = Path(__file__).parent
cases = {}


def rot(rx, ry, rz):
    R, _ = cv2.Rodrigues(np.array([rx, ry, rz], np.float64))
    return R


def add(name, K, D, R, P, size):
    Dv = np.asarray(D, np.float64).reshape(1, -1) if len(D) else None
    m1, m2 = cv2.initUndistortRectifyMap(K, Dv, R, P, size, cv2.CV_16SC2)
    cases[name] = dict(K=K, D=np.asarray(D, np.float64).reshape(-1), R=np.zeros((0, 0)) if R is None else R,
                       P=np.zeros((0, 0)) if P is None else P, size=np.array(size), map1=m1, map2=m2)


K = np.array([[1100.5, 0, 641.25], [0, 1098.75, 510.5], [0, 0, 1]], np.float64)
P = np.array([[1050.0, 0, 655.5, 0], [0, 1050.0, 500.25, 0], [0, 0, 1, 0]], np.float64)
add("d5_R_P34", K, [-0.081, 0.027, 0.0011, -0.0007, 0.013], rot(0.004, -0.006, 0.002), P, (640, 360))
add("d4_noR_noP", K, [-0.12, 0.05, 0.0, 0.0], None, None, (333, 211))
add("d8_rational", K, [0.4, -0.2, 0.001, 0.002, 0.03, 0.45, -0.15, 0.01], rot(-0.01, 0.012, 0.0), P[:, :3].copy(), (480, 270))
add("d12_prism", K, [-0.05, 0.01, 0.0005, 0.0003, 0.002, 0, 0, 0, 0.001, -0.0004, 0.0007, 0.0002], rot(0, 0.003, 0.001), P, (320, 240))
add("d14_tilt", K, [-0.05, 0.01, 0.0005, 0.0003, 0.002, 0, 0, 0, 0.001, -0.0004, 0.0007, 0.0002, 0.01, -0.02], None, P, (320, 240))
add("d0", K, [], rot(0.02, 0.01, -0.03), P, (257, 129))
flat = {}
for n, c in cases.items():
    for k, v in c.items():
        flat[f"{n}__{k}"] = v
np.savez_compressed(here / "rectmap_cv_golden.npz", cv_version=cv2.__version__, **flat)

# a calibration file in the reference's own format (keys of StereoParams::loadYAMLFile, stereo_utils.cpp:204-217)
Kr = K.copy(); Kr[0, 2] += 3.5
R = rot(0.001, 0.02, -0.002)
T = np.array([[-0.12], [0.0005], [0.001]])
Dl = np.array([[-0.081, 0.027, 0.0011, -0.0007, 0.013]])
Dr = np.array([[-0.079, 0.031, -0.0004, 0.0009, 0.011]])
size = (640, 360)
R1, R2, P1, P2, Q, _, _ = cv2.stereoRectify(K, Dl, Kr, Dr, size, R, T, flags=cv2.CALIB_ZERO_DISPARITY, alpha=0)
fs = cv2.FileStorage(str(here / "stereo_calib.yml"), cv2.FILE_STORAGE_WRITE)
for k, v in dict(leftK=K, leftD=Dl, rightK=Kr, rightD=Dr, E=np.eye(3), F=np.eye(3), R=R, T=T, R1=R1, R2=R2, P1=P1, P2=P2, Q=Q).items():
    fs.write(k, v)
fs.release()
with open(here / "stereo_calib.yml", "a") as f:  # cv::FileStorage << cv::Size writes a flow sequence
    f.write("imgsz: [ %d, %d ]\n" % size)
m = [cv2.initUndistortRectifyMap(K, Dl, R1, P1, size, cv2.CV_16SC2), cv2.initUndistortRectifyMap(Kr, Dr, R2, P2, size, cv2.CV_16SC2)]
np.savez_compressed(here / "stereo_calib_maps.npz", map00=m[0][0], map01=m[0][1], map10=m[1][0], map11=m[1][1], Q=Q)
print("cv2", cv2.__version__, "written:", list(cases))
