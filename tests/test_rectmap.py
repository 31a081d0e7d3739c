"""Row f2: rectify-map generation (EpipolarRectifyMap::compute -> cv::initUndistortRectifyMap, CV_16SC2) and the
StereoParams YAML loader.  Golden maps come from cv2 4.13 (tests/golden/make_rectmap_golden.py)."""
from pathlib import Path

import numpy as np
import pytest

GOLD = Path(__file__).parent / "golden"


def _cases():
    z = np.load(GOLD / "rectmap_cv_golden.npz")
    names = sorted({k.split("__")[0] for k in z.files if "__" in k})
    for n in names:
        yield n, {k.split("__")[1]: z[k] for k in z.files if k.startswith(n + "__")}


def test_yaml_reader_parses_the_reference_format():
    from tea_stereo_matching_b200.rectify import _read_opencv_yaml

    y = _read_opencv_yaml(str(GOLD / "stereo_calib.yml"))
    for k in ("leftK", "leftD", "rightK", "rightD", "E", "F", "R", "T", "R1", "R2", "P1", "P2", "Q", "imgsz"):
        assert k in y, k
    assert y["leftK"].shape == (3, 3) and y["P1"].shape == (3, 4) and y["Q"].shape == (4, 4) and y["leftD"].shape == (1, 5)
    assert y["leftK"][0, 0] == 1100.5 and list(y["imgsz"]) == [640.0, 360.0]
    assert y["leftD"][0, 0] == -0.081  # 17 significant digits round-trip


def test_stereo_params_errors_like_the_reference(tmp_path):
    import tea_stereo_matching_b200 as t

    with pytest.raises(ValueError):
        t.StereoParams().loadYAMLFile("")
    with pytest.raises(RuntimeError):
        t.StereoParams().loadYAMLFile(str(tmp_path / "missing.yml"))
    assert t.StereoParams().empty()


@pytest.mark.gpu
def test_maps_equal_cv2_for_every_distortion_model():
    import tea_stereo_matching_b200 as t

    for name, c in _cases():
        R = c["R"] if c["R"].size else None
        P = c["P"] if c["P"].size else None
        m1, m2 = t.initUndistortRectifyMap(c["K"], c["D"], R, P, tuple(int(v) for v in c["size"]))
        assert m1.shape == c["map1"].shape and m2.shape == c["map2"].shape, name
        bad = int((m1 != c["map1"]).any(axis=2).sum() + (m2 != c["map2"]).sum())
        assert bad == 0, f"{name}: {bad} map entries differ from cv2"


@pytest.mark.gpu
def test_stereo_params_yaml_to_maps_to_rectified_pair():
    import tea_stereo_matching_b200 as t

    sp = t.StereoParams(str(GOLD / "stereo_calib.yml"))
    want = np.load(GOLD / "stereo_calib_maps.npz")
    assert not sp.empty() and sp.imgsz == (640, 360)
    for k in ("map00", "map01", "map10", "map11"):
        assert np.array_equal(getattr(sp.map, k), want[k]), k
    Q = want["Q"]
    assert sp.rectified_f == float(np.float32(Q[2, 3])) and sp.baseline == float(np.float32(1.0) / np.float32(Q[3, 2]))
    # the maps drive the rectifier like maps loaded from OpenCV would
    rng = np.random.default_rng(2)
    frame = rng.integers(0, 256, (360, 1280, 3)).astype(np.uint8)
    r1 = t.EpipolarRectify(sp.map, sp.imgsz)
    ref = t.EpipolarRectifyMap(map00=want["map00"], map01=want["map01"], map10=want["map10"], map11=want["map11"])
    r2 = t.EpipolarRectify(ref, sp.imgsz)
    a, b = r1.rectifyStereo(frame), r2.rectifyStereo(frame)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
