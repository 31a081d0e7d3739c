"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/tsm.h
declares, validates arguments like the reference, and fails LOUDLY without a GPU."""
import ctypes as C
import re

import numpy as np
import pytest

from conftest import ROOT, _have_gpu


def test_exports_match_header(native_lib):
    from tea_stereo_matching_b200 import _native as N

    header = (ROOT / "include" / "tsm.h").read_text()
    declared = set(re.findall(r"\b(tsm_[a-z_0-9]+)\s*\(", header))
    declared -= {"tsm_ctx"}
    assert declared == set(N.EXPORTS), declared ^ set(N.EXPORTS)
    for s in declared:
        assert hasattr(native_lib, s), s
    assert native_lib.tsm_version() == 100


def test_status_strings(native_lib):
    assert native_lib.tsm_status_string(0) == b"ok"
    assert b"unsupported" in native_lib.tsm_status_string(4)


@pytest.mark.skipif(_have_gpu(), reason="only meaningful without a GPU")
def test_no_cpu_fallback_without_gpu(native_lib):
    import tea_stereo_matching_b200 as t

    with pytest.raises(t.ADCensusError) as e:
        t.Context(0)
    assert "no CPU fallback" in str(e.value)
    m = t.ADCensus()
    m.setMatchingStrategy(t.ColorModel.RGB)
    m.setMinMaxDisparity(0, 16)
    img = np.zeros((16, 16, 3), np.uint8)
    with pytest.raises(t.ADCensusError):
        m.compute(img, img)


def test_setter_errors_mirror_reference():
    import tea_stereo_matching_b200 as t

    m = t.ADCensus()
    assert (m._min, m._max, m._model) == (0, 64, t.ColorModel.HSI)  # ADCensus.cpp:409-420
    with pytest.raises(t.ADCensusError, match="Set MinMaxDisparity error"):
        m.setMinMaxDisparity(5, 5)
    with pytest.raises(t.ADCensusError, match="Set MinMaxDisparity error"):
        m.setMinMaxDisparity(-1, 5)
    with pytest.raises(t.ADCensusError, match="Offset must be positive"):
        m.setOffset(-1)
    with pytest.raises(t.ADCensusError, match="Image error"):
        m.compute(np.zeros((0, 0, 3), np.uint8), np.zeros((0, 0, 3), np.uint8))
    with pytest.raises(t.ADCensusError, match="Image error"):
        m.compute(np.zeros((8, 8, 3), np.uint8), np.zeros((8, 9, 3), np.uint8))


def test_rectify_mirror_error_behaviour(capsys):
    import tea_stereo_matching_b200 as t

    r = t.EpipolarRectify()
    assert r.rectify(np.zeros((4, 8, 3), np.uint8)) is None  # logs + returns, EpipolarRectify.cpp:70-79
    assert "params is empty" in capsys.readouterr().err
    with pytest.raises(RuntimeError, match="stereo params is empty"):
        r.loadEpipolarRectifyMap(t.EpipolarRectifyMap(), (8, 4))
