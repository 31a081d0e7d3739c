#!/usr/bin/env python
"""Makes tests/golden/consumers_cv_golden.npz: outputs of the two OpenCV calls inside the reference's
reprojectTo3D(disparity, Q, XYZ) (source/stereo.cpp:192-198) -- cv::gemm on CV_32F and cv::divide -- from cv2."""
from pathlib import Path

import cv2
import numpy as np

rng = np.random.default_rng(7)
H, W = 37, 53
disp = rng.uniform(-2, 200, (H, W)).astype(np.float32)
disp[rng.random((H, W)) < 0.1] = -1
disp[3, 4] = 0.0
Q = np.array([[1, 0, 0, -640.25], [0, 1, 0, -511.5], [0, 0, 0, 1100.125], [0, 0, 1 / 0.12, 0.3]], np.float64)
u = np.broadcast_to(np.arange(W, dtype=np.float32)[None, :], (H, W)).reshape(1, -1)
v = np.broadcast_to(np.arange(H, dtype=np.float32)[:, None], (H, W)).reshape(1, -1)
pix = np.concatenate([u, v, disp.reshape(1, -1), np.ones((1, H * W), np.float32)], 0)
xyzw = cv2.gemm(Q.astype(np.float32), pix, 1.0, None, 0.0)
rows = [cv2.divide(xyzw[i : i + 1], xyzw[3:4]) for i in range(3)]
np.savez_compressed(Path(__file__).parent / "consumers_cv_golden.npz", disp=disp, Q=Q, xyzw=xyzw, xyz=np.concatenate(rows, 0),
                    cv_version=cv2.__version__)
print("cv2", cv2.__version__, "written")
