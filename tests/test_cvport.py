"""oracle/cvport.c (integer restatements of the five OpenCV calls on the path) against
cv2 4.13.0: committed golden vectors always, live cv2 on random inputs when importable."""
import ctypes as C

import numpy as np
import pytest

from conftest import GOLDEN


def P(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.fixture(scope="module")
def g():
    return np.load(GOLDEN / "cv_golden.npz")


def test_equalize_blur_canny_golden(port, g):
    L = port.lib
    for n in "abc":
        img = np.ascontiguousarray(g[f"u8_{n}"])
        H, W = img.shape
        out = np.empty_like(img)
        L.cvp_equalize_hist(P(img), P(out), H, W)
        assert np.array_equal(out, g[f"eq_{n}"])
        L.cvp_blur3x3(P(img), P(out), H, W)
        assert np.array_equal(out, g[f"blur_{n}"])
        L.cvp_canny3(P(img), P(out), H, W, 30, 90)
        assert np.array_equal(out, g[f"canny_{n}"])


def test_median_golden(port, g):
    f = np.ascontiguousarray(g["f32"])
    out = np.empty_like(f)
    port.lib.cvp_median3x3_f32(P(f), P(out), *f.shape)
    assert np.array_equal(out, g["median"])


def test_remap_and_convertmaps_golden(port, g):
    src = np.ascontiguousarray(g["remap_src"])
    mx, my = np.ascontiguousarray(g["remap_mx"]), np.ascontiguousarray(g["remap_my"])
    H, W = mx.shape
    m1 = np.empty((H, W, 2), np.int16)
    m2 = np.empty((H, W), np.uint16)
    port.lib.cvp_convert_maps_f32(P(mx), P(my), H, W, P(m1), P(m2))
    assert np.array_equal(m1, g["remap_m1"]) and np.array_equal(m2, g["remap_m2"])
    out = np.empty((H, W, 3), np.uint8)
    port.lib.cvp_remap_bilinear_8uc3_fixed(P(src), src.shape[0], src.shape[1], src.strides[0], P(m1), P(m2), P(out), H, W)
    assert np.array_equal(out, g["remap_fixed"])
    assert np.array_equal(out, g["remap_float"])  # float maps quantise to the same 1/32 grid


def test_live_cv2_random(port):
    cv2 = pytest.importorskip("cv2")
    if not cv2.__version__.startswith("4.13"):
        pytest.skip("cv2 is not 4.13")
    L = port.lib
    rng = np.random.default_rng(11)
    for t in range(12):
        H, W = int(rng.integers(3, 70)), int(rng.integers(3, 90))
        img = rng.integers(0, 256, (H, W), dtype=np.uint8) if t % 2 else cv2.GaussianBlur(
            rng.integers(0, 256, (H, W)).astype(np.float32), (0, 0), 2.5).astype(np.uint8)
        out = np.empty_like(img)
        L.cvp_equalize_hist(P(img), P(out), H, W)
        assert np.array_equal(out, cv2.equalizeHist(img))
        L.cvp_blur3x3(P(img), P(out), H, W)
        assert np.array_equal(out, cv2.blur(img, (3, 3)))
        L.cvp_canny3(P(img), P(out), H, W, 30, 90)
        assert np.array_equal(out, cv2.Canny(img, 30, 90, apertureSize=3))
        f = rng.normal(0, 9, (H, W)).astype(np.float32)
        fo = np.empty_like(f)
        L.cvp_median3x3_f32(P(f), P(fo), H, W)
        assert np.array_equal(fo, cv2.medianBlur(f, 3))
