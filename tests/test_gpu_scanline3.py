"""The blocked scanline walk (k_scanline3.cu) against the striped one (k_scanline.cu, itself bit-exact against the reference's
golden vectors in test_gpu_parity.py): same volumes, same WTA maps, same final maps, bit for bit, over the geometry corners the
two kernels treat differently -- odd widths (the 16-byte tail chunk of two pixels changes parity from row to row), narrow and
wide tails, every vector width of the blocked loads, a non-zero minimum disparity, mask matching, partial last CTAs.
TSM_SCAN3=0 selects the striped kernel; the switch is read per call."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

# H, W, minD, maxD  ->  Dn = maxD - minD + 1 = 32 * KM + r
GEOMETRIES = [
    (97, 161, 0, 40),    # KM 1, r 9  (tail pitch 16), odd width and height
    (60, 200, 0, 33),    # KM 1, r 2  (tail pitch 2: two pixels per 16-byte chunk)
    (64, 151, 0, 64),    # KM 2, r 1, odd width: the chunk offset toggles along a vertical path
    (50, 330, 0, 100),   # KM 3, r 5  (tail pitch 8), scalar blocked loads
    (40, 300, 3, 131),   # KM 4 (128-bit loads), r 1, minD 3
    (48, 420, 10, 200),  # KM 5, r 31 (tail pitch 32), minD 10
    (33, 500, 0, 192),   # KM 6, r 1: the benchmark's range
    (30, 700, 0, 250),   # KM 7, r 27
    (36, 600, 0, 280),   # KM 8, r 25
    (24, 900, 0, 384),   # KM 12, r 1
    (2, 4, 0, 33),       # the smallest pair the blocked walk takes: one step per vertical path, one partial CTA
    (3, 5, 0, 40),       # fewer columns than a CTA owns, odd width
    (7, 9, 0, 33),       # a single partial pixel group per row
    (300, 13, 0, 64),    # tall and narrow: one CTA per view in the vertical launch
]


def _both(fn):
    out = []
    for flag in ("0", "1"):
        os.environ["TSM_SCAN3"] = flag
        try:
            out.append(fn())
        finally:
            os.environ.pop("TSM_SCAN3", None)
    return out


@pytest.mark.parametrize("H,W,mind,maxd", GEOMETRIES)
def test_blocked_walk_equals_striped_walk_stage_by_stage(H, W, mind, maxd):
    import tea_stereo_matching_b200 as t
    from tea_stereo_matching_b200 import _native as N
    from tea_stereo_matching_b200.synth import synth_v1

    if min(H, W) < 16:  # too small for the synthetic scene generator: smooth noise (similar neighbours do occur)
        rng = np.random.default_rng(H * 1000 + W)
        left = (rng.integers(0, 4, (H, W, 3)) * 8 + 100).astype(np.uint8)
        right = (rng.integers(0, 4, (H, W, 3)) * 8 + 100).astype(np.uint8)
    else:
        left, right = synth_v1(H, W, min(maxd, W // 3), seed=500 + W)

    def run():
        r = t.StageRunner(left, right, maxd, min_disparity=mind)
        r.run(N.STAGE_PREP | N.STAGE_INIT | N.STAGE_AGGREGATE | N.STAGE_SCANLINE)
        res = [r.volume(0), r.volume(1), r.wta(0), r.wta(1)]
        r.close()
        return res

    a, b = _both(run)
    for k, name in enumerate(("left volume", "right volume", "left WTA", "right WTA")):
        assert np.array_equal(a[k], b[k]), (name, int((a[k] != b[k]).sum()))


def test_blocked_walk_equals_striped_walk_with_mask_and_roi():
    import tea_stereo_matching_b200 as t
    from tea_stereo_matching_b200.synth import synth_v1

    left, right = synth_v1(96, 260, 60, seed=77)
    left, right = left.copy(), right.copy()
    left[20:50, 40:90] = 0    # holes: black pixels of both views
    right[25:60, 30:70] = 0
    left[70:75, :] = 0
    for model, roi, mask in ((t.ColorModel.RGB, False, True), (t.ColorModel.HSI, False, True), (t.ColorModel.RGB, True, True)):
        def run():
            m = t.ADCensus()
            m.setMatchingStrategy(model, roi, mask)
            m.setMinMaxDisparity(0, 70)  # ROI mode: replaced by W / 2 = 130 -> KM 4, r 3
            return m.compute(left, right)

        a, b = _both(run)
        assert np.array_equal(a, b), (model, roi, mask, int((a != b).sum()))


def test_odd_width_against_the_reference_live(ref):
    """Odd width, one tail element (two pixels per 16-byte tail chunk, alternating offset): both walks against the compiled
    reference, fed with the reference's aggregated volumes."""
    import tea_stereo_matching_b200 as t
    from tea_stereo_matching_b200 import _native as N
    from tea_stereo_matching_b200.synth import synth_v1

    H, W, D = 64, 151, 64
    left, right = synth_v1(H, W, W // 3, seed=500 + W)
    st = ref.run(left, right, D, serial_scanline=True)

    def run():
        r = t.StageRunner(left, right, D)
        r.run(N.STAGE_PREP)
        for v in range(2):
            r.set_volume(v, st.vol_agg[v])
        r.run(N.STAGE_SCANLINE)
        res = [r.volume(0), r.volume(1), r.wta(0), r.wta(1)]
        r.close()
        return res

    for res in _both(run):
        for v in range(2):
            assert np.array_equal(res[v], st.vol_scan[v]), ("volume", v)
            assert np.array_equal(res[2 + v], st.wta[v].astype(np.int32)), ("wta", v)
