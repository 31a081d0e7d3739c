"""The C++20 facade (tea_stereo_matching_b200/cpp): same class names / signatures / exception
types as the reference's stereo::ADCensus and stereo::EpipolarRectify."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

CPP = ROOT / "tea_stereo_matching_b200" / "cpp"
DEMO = ROOT / "tests" / "cpp" / "facade_demo"


@pytest.fixture(scope="module")
def demo(native_lib):
    env = {k: v for k, v in os.environ.items() if k not in ("CXX", "CC")}
    subprocess.run(["make", "-s", "-C", str(CPP)], check=True, env=env)
    subprocess.run(["g++", "-std=c++20", "-O1", "-o", str(DEMO), str(ROOT / "tests" / "cpp" / "facade_demo.cpp"),
                    f"-L{ROOT / 'tea_stereo_matching_b200'}", "-ltea_stereo", "-ltsm_b200",
                    f"-Wl,-rpath,{ROOT / 'tea_stereo_matching_b200'}"], check=True, env=env)
    return DEMO


def test_facade_api_error_behaviour(demo):
    r = subprocess.run([str(demo), "api"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "API CHECKS PASSED" in r.stdout
    assert "Set MinMaxDisparity error" in r.stdout and "Image error" in r.stdout


@pytest.mark.gpu
def test_facade_compute_matches_python_mirror_and_oracle(demo, pair_0600, port, tmp_path):
    import tea_stereo_matching_b200 as t

    left, right = pair_0600
    H, W, _ = left.shape
    D = 48
    inp, out = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(inp, "wb") as f:
        f.write(np.array([H, W, D], np.int32).tobytes())
        f.write(left.tobytes())
        f.write(right.tobytes())
    r = subprocess.run([str(demo), "run", str(inp), str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = np.fromfile(out, np.float32).reshape(H, W)
    m = t.ADCensus()
    m.setMatchingStrategy(t.ColorModel.RGB)
    m.setMinMaxDisparity(0, D)
    assert np.array_equal(got, m.compute(left, right))  # same kernels behind both host mirrors
    want = port.compute(left, right, D)
    diff = np.abs(got.astype(np.float64) - want)
    assert (diff > 1).mean() <= 1e-3 and (diff > 0.05).mean() <= 1e-3
    r = subprocess.run([str(demo), "batch", str(inp), str(out), "3"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    b = np.fromfile(out, np.float32).reshape(3, H, W)
    assert all(np.array_equal(b[i], got) for i in range(3))
    # setDevice(-1): the same batch sharded over every visible device (thread per GPU, four pairs in flight each) returns the same bits
    r = subprocess.run([str(demo), "batchall", str(inp), str(out), "7"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    b = np.fromfile(out, np.float32).reshape(7, H, W)
    assert all(np.array_equal(b[i], got) for i in range(7))


@pytest.mark.gpu
def test_facade_consumers_match_oracle(demo, pair_0600, tmp_path):
    from oracle import consumers_oracle as co

    left, right = pair_0600
    H, W, _ = left.shape
    inp, out = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(inp, "wb") as f:
        f.write(np.array([H, W, 48], np.int32).tobytes())
        f.write(left.tobytes())
        f.write(right.tobytes())
    r = subprocess.run([str(demo), "consume", str(inp), str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    raw = np.fromfile(out, np.uint8)
    n = H * W
    disp = raw[: 4 * n].view(np.float32).reshape(H, W)
    depth = raw[4 * n : 8 * n].view(np.float32).reshape(H, W)
    xyz = raw[8 * n : 20 * n].view(np.float32).reshape(H, W, 3)
    xyzq = raw[20 * n : 32 * n].view(np.float32).reshape(H, W, 3)
    color = raw[32 * n : 35 * n].reshape(H, W, 3)
    Q = np.array([[1, 0, 0, -W / 2.0], [0, 1, 0, -H / 2.0], [0, 0, 0, 700.0], [0, 0, 10.0, 0.5]])
    eq = lambda a, b: np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert eq(depth, co.reproject_to_depth(disp, 700.0, 0.1))
    assert eq(xyz, co.reproject_to_3d(disp, 700.0, 0.1, W / 2.0, H / 2.0))
    assert eq(xyzq, co.reproject_to_3d_q(disp, Q))
    assert np.array_equal(color, co.apply_colormap(disp, co.jet_colormap()))


@pytest.mark.gpu
def test_facade_stereo_params_yaml_to_maps(demo, tmp_path):
    gold = ROOT / "tests" / "golden"
    out = tmp_path / "maps.bin"
    r = subprocess.run([str(demo), "params", str(gold / "stereo_calib.yml"), str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    want = np.load(gold / "stereo_calib_maps.npz")
    H, W = want["map01"].shape
    raw = np.fromfile(out, np.uint8)
    n = H * W
    o = 0
    for k, nb, dt, shape in (("map00", 4, np.int16, (H, W, 2)), ("map01", 2, np.uint16, (H, W)), ("map10", 4, np.int16, (H, W, 2)),
                             ("map11", 2, np.uint16, (H, W))):
        got = raw[o : o + n * nb].view(dt).reshape(shape)
        assert np.array_equal(got, want[k]), k
        o += n * nb
    f, cx, cy, B = raw[o : o + 16].view(np.float32)
    Q = want["Q"]
    assert f == np.float32(Q[2, 3]) and cx == np.float32(-Q[0, 3]) and cy == np.float32(-Q[1, 3])
    assert B == np.float32(1.0) / np.float32(Q[3, 2])
