"""The oracle port (oracle/adcensus_oracle.c) against the reference-generated golden
vectors (tests/golden/*.npz, made by tests/golden/make_golden.py from the UNMODIFIED
reference ADCensus.cpp, serial-scanline semantics).  Everything must be bit-exact."""
import hashlib

import numpy as np
import pytest


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def check_against_golden(st, g, full_volumes):
    for name in ("vol_init", "vol_agg", "vol_scan"):
        for k in range(2):
            v = getattr(st, name)[k]
            assert np.array_equal(v[::7, ::5, :], g[f"{name}{k}_sample"]), f"{name}{k} sample"
            assert sha(v) == str(g[f"{name}{k}_sha256"]), f"{name}{k} sha256"
            if full_volumes:
                assert np.array_equal(v, g[f"{name}{k}"])
    for k in range(2):
        assert np.array_equal(np.stack(st.arms[k], axis=2), g[f"arms{k}"]), f"arms{k}"
        assert np.array_equal(st.wta[k], g[f"wta{k}"]), f"wta{k}"
    assert np.array_equal(st.lrc, g["lrc"])
    for i in range(5):
        assert np.array_equal(st.vote[i], g[f"vote{i}"]), f"vote{i}"
    assert np.array_equal(st.interp, g["interp"])
    assert np.array_equal(st.discont, g["discont"])
    assert np.array_equal(st.final, g["final"])


def test_port_matches_reference_golden_0600(port, pair_0600, golden_0600):
    left, right = pair_0600
    st = port.run(left, right, int(golden_0600["max_disparity"]))
    check_against_golden(st, golden_0600, False)


def test_port_matches_reference_golden_synth(port, golden_synth):
    st = port.run(golden_synth["left"], golden_synth["right"], int(golden_synth["max_disparity"]))
    check_against_golden(st, golden_synth, True)


def test_census_signature_form_equals_direct_form(port, pair_0600):
    """popc((ltL&gtR)|(gtL&ltR)) summed over channels == the literal sign-product count (ADCensus.cpp:461-472)."""
    import ctypes as C

    left, right = pair_0600
    H, W, _ = left.shape
    lib = port.lib
    lib.orc_census_direct.restype = C.c_int
    lt = [np.zeros((H, W, 3), np.uint64) for _ in range(2)]
    gt = [np.zeros((H, W, 3), np.uint64) for _ in range(2)]
    for img, l, g in ((left, lt[0], gt[0]), (right, lt[1], gt[1])):
        lib.orc_census_signatures(img.ctypes.data_as(C.c_void_p), H, W, l.ctypes.data_as(C.c_void_p), g.ctypes.data_as(C.c_void_p))
    rng = np.random.default_rng(3)
    for _ in range(2000):
        y = int(rng.integers(3, H - 3))
        xl = int(rng.integers(4, W - 4))
        xr = int(rng.integers(4, W - 4))
        direct = lib.orc_census_direct(left.ctypes.data_as(C.c_void_p), right.ctypes.data_as(C.c_void_p), W, y, xl, xr)
        sig = 0
        for c in range(3):
            v = (int(lt[0][y, xl, c]) & int(gt[1][y, xr, c])) | (int(gt[0][y, xl, c]) & int(lt[1][y, xr, c]))
            sig += bin(v).count("1")
        assert sig == direct


def test_live_reference_if_present(port, ref, pair_0600):
    """Where the reference .so exists, re-derive the golden comparison live, incl. the racy/serial distinction."""
    left, right = pair_0600
    left, right = left[:120, :200].copy(), right[:120, :200].copy()
    a = port.run(left, right, 32)
    b = ref.run(left, right, 32, serial_scanline=True)
    for name in ("vol_init", "vol_agg", "vol_scan"):
        for k in range(2):
            assert np.array_equal(getattr(a, name)[k], getattr(b, name)[k]), (name, k)
    assert np.array_equal(a.final, b.final)
    # integer AD sums / census counts straight from the reference's own functions
    rng = np.random.default_rng(5)
    n = 500
    H, W, _ = left.shape
    y = rng.integers(3, H - 3, n)
    xl = rng.integers(4, W - 4, n)
    xr = rng.integers(4, W - 4, n)
    ad3, cen, cost = ref.ad_census_pairs(left, right, y, xl, xr)
    import ctypes as C

    port.lib.orc_ad3.restype = C.c_int
    port.lib.orc_census_direct.restype = C.c_int
    for i in range(n):
        assert port.lib.orc_ad3(left.ctypes.data_as(C.c_void_p), right.ctypes.data_as(C.c_void_p), W, int(y[i]), int(xl[i]), int(xr[i])) == ad3[i]
        assert port.lib.orc_census_direct(left.ctypes.data_as(C.c_void_p), right.ctypes.data_as(C.c_void_p), W, int(y[i]), int(xl[i]), int(xr[i])) == cen[i]


def test_reference_compute_entry_equals_staged(ref, golden_synth):
    """ADCensus::compute (public entry) == our staged call sequence when both run single-threaded."""
    import os
    import subprocess
    import sys

    code = (
        "import numpy as np, oracle, sys;"
        "g=np.load(sys.argv[1]); r=oracle.Ref();"
        "out,_=r.compute(g['left'],g['right'],int(g['max_disparity']));"
        "assert np.array_equal(out,g['final']); print('ok')"
    )
    from conftest import GOLDEN, ROOT

    env = dict(os.environ, OMP_NUM_THREADS="1", PYTHONPATH=str(ROOT))
    r = subprocess.run([sys.executable, "-c", code, str(GOLDEN / "ref_synth_96x128_d24.npz")], env=env, capture_output=True, text=True)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr
