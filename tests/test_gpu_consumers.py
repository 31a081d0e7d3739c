"""Disparity consumers through the C-ABI against the CPU restatement: bit-exact."""
import numpy as np
import pytest

from oracle import consumers_oracle as co

pytestmark = pytest.mark.gpu


def _maps():
    rng = np.random.default_rng(5)
    d = rng.uniform(0, 192, (211, 333)).astype(np.float32)
    d[rng.random(d.shape) < 0.15] = -1
    d[rng.random(d.shape) < 0.05] = -2
    d[0, :7] = [0.0, -0.0, np.inf, -np.inf, np.nan, 1e-30, 3e38]
    yield "random+specials", d
    yield "all invalid", np.full((9, 17), -1, np.float32)
    yield "constant", np.full((16, 16), 12.25, np.float32)


def _same(a, b):
    """Bit patterns equal; NaNs only have to be NaNs at the same places (the payload of a propagated NaN is
    hardware-specific: x86 keeps the operand's, the GPU returns the canonical one)."""
    if a.shape != b.shape:
        return False
    if a.dtype != np.float32:
        return np.array_equal(a, b)
    na, nb = np.isnan(a), np.isnan(b)
    return np.array_equal(na, nb) and np.array_equal(a.view(np.uint32)[~na], b.view(np.uint32)[~nb])


def test_reproject_and_colormap_match_oracle_bit_for_bit():
    import tea_stereo_matching_b200 as t

    Q = np.array([[1, 0, 0, -166.5], [0, 1, 0, -105.25], [0, 0, 0, 1100.5], [0, 0, 1 / 0.119, 0.25]])
    assert np.array_equal(t.JETColorMap(), co.jet_colormap())
    other = np.random.default_rng(3).integers(0, 256, (1, 256, 3)).astype(np.uint8)
    for name, d in _maps():
        assert _same(t.reprojectToDepth(d, 1100.5, 0.119), co.reproject_to_depth(d, 1100.5, 0.119)), name
        assert _same(t.reprojectTo3D(d, 1100.5, 0.119, 166.5, 105.25), co.reproject_to_3d(d, 1100.5, 0.119, 166.5, 105.25)), name
        assert _same(t.reprojectTo3D(d, Q), co.reproject_to_3d_q(d, Q)), name
        assert _same(t.applyColorMap(d), co.apply_colormap(d, co.jet_colormap())), name
        assert _same(t.applyColorMap(d, 10.0, 150.0), co.apply_colormap(d, co.jet_colormap(), 10.0, 150.0)), name
        assert _same(t.applyColorMap(d, colorMap=other), co.apply_colormap(d, other)), name


def test_consuming_the_matchers_last_map_on_the_device(pair_0600):
    import tea_stereo_matching_b200 as t

    left, right = pair_0600
    m = t.ADCensus()
    m.setMatchingStrategy(t.ColorModel.RGB)
    m.setMinMaxDisparity(0, 48)
    disp = m.compute(left, right)
    assert _same(t.reprojectToDepth(m, 700.0, 0.1), co.reproject_to_depth(disp, 700.0, 0.1))
    assert _same(t.reprojectTo3D(m, 700.0, 0.1, 160.0, 90.0), co.reproject_to_3d(disp, 700.0, 0.1, 160.0, 90.0))
    assert _same(t.applyColorMap(m), co.apply_colormap(disp, co.jet_colormap()))
    m2 = t.ADCensus()
    with pytest.raises(t.ADCensusError):
        t.reprojectToDepth(m2, 700.0, 0.1)  # nothing computed on that context yet
