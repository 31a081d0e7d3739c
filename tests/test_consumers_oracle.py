"""Pins oracle/consumers_oracle.py (CPU restatement of source/stereo.cpp:75-202) against OpenCV 4.13 for the two
OpenCV calls the reference makes (cv::gemm on CV_32F, cv::divide): committed vectors, and live where cv2 imports."""
from pathlib import Path

import numpy as np
import pytest

from oracle import consumers_oracle as co

GOLD = Path(__file__).parent / "golden" / "consumers_cv_golden.npz"


def _pixels(disp):
    H, W = disp.shape
    u = np.broadcast_to(np.arange(W, dtype=np.float32)[None, :], (H, W)).reshape(1, -1)
    v = np.broadcast_to(np.arange(H, dtype=np.float32)[:, None], (H, W)).reshape(1, -1)
    return np.concatenate([u, v, disp.reshape(1, -1), np.ones((1, H * W), np.float32)], 0)


def test_gemm_and_divide_match_committed_cv2_vectors():
    z = np.load(GOLD)
    disp, Q = z["disp"], z["Q"]
    xyzw = co.gemm32f(Q.astype(np.float32), _pixels(disp))
    assert np.array_equal(xyzw.view(np.uint32), z["xyzw"].view(np.uint32))
    got = co.reproject_to_3d_q(disp, Q)
    want = z["xyz"].T.reshape(disp.shape + (3,))
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))  # bit pattern: inf / nan included


def test_gemm_and_divide_match_live_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(11)
    disp = rng.uniform(-3, 300, (64, 80)).astype(np.float32)
    Q = rng.normal(0, 50, (4, 4))
    pix = _pixels(disp)
    xyzw = cv2.gemm(Q.astype(np.float32), pix, 1.0, None, 0.0)
    assert np.array_equal(co.gemm32f(Q.astype(np.float32), pix).view(np.uint32), xyzw.view(np.uint32))
    want = np.concatenate([cv2.divide(xyzw[i : i + 1], xyzw[3:4]) for i in range(3)], 0).T.reshape(64, 80, 3)
    assert np.array_equal(co.reproject_to_3d_q(disp, Q).view(np.uint32), want.view(np.uint32))


def test_jet_table_shape_and_anchors():
    t = co.jet_colormap()
    assert t.shape == (1, 256, 3) and t.dtype == np.uint8
    assert tuple(t[0, 0]) == (128, 0, 0) and tuple(t[0, 32]) == (255, 0, 0) and tuple(t[0, 96]) == (254, 255, 2)
    assert tuple(t[0, 159]) == (1, 255, 254) and tuple(t[0, 255]) == (0, 0, 128)


def test_depth_and_xyz_keep_invalid_pixels_zero():
    d = np.array([[-1, -2, 0, 4], [np.inf, 16, 0.5, -0.0]], np.float32)
    depth = co.reproject_to_depth(d, 1000.0, 0.12)
    assert depth[0, 0] == 0 and depth[0, 1] == 0 and depth[1, 0] == 0
    assert np.isinf(depth[0, 2]) and depth[0, 3] == np.float32(np.float32(1000.0) * np.float32(0.12)) / np.float32(4)
    xyz = co.reproject_to_3d(d, 1000.0, 0.12, 1.5, 0.5)
    assert np.all(xyz[0, 0] == 0) and np.all(xyz[1, 0] == 0) and xyz[1, 1, 2] == depth[1, 1]


def test_colormap_auto_and_explicit_range():
    d = np.array([[-1, 0, 5, 10], [2.5, np.nan, 7.5, -2]], np.float32)
    cm = co.jet_colormap()
    a = co.apply_colormap(d, cm)
    assert np.all(a[0, 0] == 0) and np.all(a[1, 3] == 0)
    assert np.array_equal(a[0, 1], cm[0, 0]) and np.array_equal(a[0, 3], cm[0, 255]) and np.array_equal(a[0, 2], cm[0, 127])
    b = co.apply_colormap(d, cm, 2.0, 8.0)
    assert np.all(b[0, 1] == 0) and np.all(b[0, 3] == 0) and np.array_equal(b[1, 0], cm[0, int(np.float32(0.5 / 6) * 255)])


def test_point_cloud_writers_match_the_restatement(tmp_path):
    """writePointCloudToPCD / writePointCloudToPLY (stereo.cpp:204-356) are host file I/O: the library's writers against the
    oracle's restatement of std::to_chars formatting, byte for byte (no GPU involved)."""
    import tea_stereo_matching_b200 as t
    from oracle import consumers_oracle as co

    rng = np.random.default_rng(11)
    H, W = 23, 31
    xyz = (rng.standard_normal((H, W, 3)) * rng.choice([1e-6, 1e-3, 1.0, 37.5, 1e4, 1e9], (H, W, 1))).astype(np.float32)
    xyz[3, 4] = (np.inf, 1.0, 2.0)       # dropped
    xyz[5, 6] = (1.0, 2.0, np.inf)       # dropped
    xyz[7, 8] = (0.0, -0.0, 100000.0)    # zeros and an integer-valued float
    xyz[9, 1] = (1e-5, 123456792.0, 0.1)
    xyz[2, 2] = (-np.inf, 1.0, 1.0)      # kept: only +infinity is filtered
    bgr = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    for fmt, fn in (("pcd", t.writePointCloudToPCD), ("ply", t.writePointCloudToPLY)):
        path = tmp_path / f"cloud.{fmt}"
        n = fn(bgr, xyz, str(path))
        assert n == H * W - 2
        assert path.read_bytes() == co.point_cloud_text(bgr, xyz, fmt), fmt
    assert t.writePointCloudToPCD(None, xyz, "x.pcd") == 0  # empty input: logged, nothing written (stereo.cpp:252-256)
