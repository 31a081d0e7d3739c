"""GPU parity tests: every CUDA stage, called through the C-ABI (tsm_stage_*), against
the oracle on the ORACLE's input for that stage, so errors never compound.

Bars (BASELINE.json north_star):
  census signatures, arms, integer maps, WTA ........ bit-exact
  initial cost, scanline, sub-pixel ................. bit-exact (fp32, same op order)
  aggregated cost ................................... <= 2e-6 relative (fp64 prefix sums vs the
                                                       reference's sequential fp32 sums; contract 1e-4)
  end-to-end final disparity ........................ |d| <= 0.05 px, <= 0.1 % of pixels off by > 1 px
"""
import numpy as np
import pytest

from tea_stereo_matching_b200.synth import synth_v1

pytestmark = pytest.mark.gpu

AGG_RTOL = 2e-6


def _runner(left, right, maxd):
    import tea_stereo_matching_b200 as t
    from tea_stereo_matching_b200 import _native as N

    return t.StageRunner(left, right, maxd), N


def _cases(pair_0600, golden_synth):
    left, right = pair_0600
    return [
        ("0600_320x180_d48", left, right, 48),
        ("synth_96x128_d24", golden_synth["left"], golden_synth["right"], int(golden_synth["max_disparity"])),
        ("0600_crop_ragged_d70", left[11:150, 7:300].copy(), right[11:150, 7:300].copy(), 70),
        # BASELINE disparity ranges on small images: Dn = 129 / 193 / 257 / 385 (K = 5 / 7 / 9 / 13 registers, tail = 1)
        ("synth_60x400_d128",) + synth_v1(60, 400, 128, seed=31) + (128,),
        ("synth_72x420_d192",) + synth_v1(72, 420, 192, seed=32) + (192,),
        ("synth_56x520_d256",) + synth_v1(56, 520, 256, seed=33) + (256,),
        ("synth_gray_52x700_d384",) + synth_v1(52, 700, 384, seed=34, gray=True) + (384,),
        # Dn a multiple of 32 (no tail part) and Dn < 32 (tail only)
        ("synth_64x200_d63",) + synth_v1(64, 200, 63, seed=35) + (63,),
        ("synth_50x90_d15",) + synth_v1(50, 90, 15, seed=36) + (15,),
    ]


@pytest.fixture(scope="module")
def cases(pair_0600, golden_synth, port, native_lib):
    out = []
    for name, l, r, d in _cases(pair_0600, golden_synth):
        out.append((name, l, r, d, port.run(l, r, d)))
    return out


def test_prep_census_and_arms_bit_exact(cases, port):
    import ctypes as C

    for name, l, r, d, st in cases:
        run, N = _runner(l, r, d)
        run.run(N.STAGE_PREP)
        H, W, _ = l.shape
        for view, img in enumerate((l, r)):
            arms = run.arms(view)
            assert np.array_equal(arms, np.stack(st.arms[view], axis=2)), (name, "arms", view)
            lt = np.zeros((H, W, 3), np.uint64)
            gt = np.zeros((H, W, 3), np.uint64)
            port.lib.orc_census_signatures(img.ctypes.data_as(C.c_void_p), H, W, lt.ctypes.data_as(C.c_void_p),
                                           gt.ctypes.data_as(C.c_void_p))
            cen = run.census(view)
            for c in range(3):
                assert np.array_equal(cen[c], lt[:, :, c]), (name, "lt", view, c)
                assert np.array_equal(cen[3 + c], gt[:, :, c]), (name, "gt", view, c)
        run.close()


def test_cost_init_bit_exact(cases):
    for name, l, r, d, st in cases:
        run, N = _runner(l, r, d)
        run.run(N.STAGE_PREP | N.STAGE_INIT)
        for view in range(2):
            assert np.array_equal(run.volume(view), st.vol_init[view]), (name, view)
        run.close()


def test_aggregate_within_few_ulp(cases):
    for name, l, r, d, st in cases:
        run, N = _runner(l, r, d)
        run.run(N.STAGE_PREP)
        for view in range(2):
            run.set_volume(view, st.vol_init[view])
        run.run(N.STAGE_AGGREGATE)
        for view in range(2):
            got, want = run.volume(view), st.vol_agg[view]
            rel = np.abs(got - want) / np.maximum(np.abs(want), 1e-30)
            assert rel.max() <= AGG_RTOL, (name, view, float(rel.max()))
        run.close()


def test_scanline_bit_exact(cases):
    for name, l, r, d, st in cases:
        run, N = _runner(l, r, d)
        run.run(N.STAGE_PREP)
        for view in range(2):
            run.set_volume(view, st.vol_agg[view])
        run.run(N.STAGE_SCANLINE)
        for view in range(2):
            got = run.volume(view)
            assert np.array_equal(got, st.vol_scan[view]), (name, view, int((got != st.vol_scan[view]).sum()))
            # the last pass also emits the WTA map (cost2disparity fused into the leftward pass)
            assert np.array_equal(run.wta(view), st.wta[view]), (name, "fused wta", view)
        run.close()


def test_post_chain_bit_exact_stage_by_stage(cases):
    for name, l, r, d, st in cases:
        run, N = _runner(l, r, d)
        run.run(N.STAGE_PREP)
        for view in range(2):
            run.set_volume(view, st.vol_scan[view])
        run.run(N.STAGE_WTA)
        for view in range(2):
            assert np.array_equal(run.wta(view), st.wta[view]), (name, "wta", view)
        run.run(N.STAGE_LRC)
        assert np.array_equal(run.disp(), st.lrc), (name, "lrc")
        prev = st.lrc
        for i in range(5):
            run.set_disp(prev)
            run.run(N.STAGE_VOTE, i)
            got = run.disp()
            assert np.array_equal(got, st.vote[i]), (name, "vote", i, int((got != st.vote[i]).sum()))
            prev = st.vote[i]
        run.set_disp(st.vote[4])
        run.run(N.STAGE_INTERP)
        assert np.array_equal(run.disp(), st.interp), (name, "interp")
        run.set_disp(st.interp)
        run.run(N.STAGE_DISCONT)
        assert np.array_equal(run.disp(), st.discont), (name, "discont")
        run.set_disp(st.discont)
        run.run(N.STAGE_SUBPIXEL)
        assert np.array_equal(run.final(), st.final), (name, "final")
        run.close()


def test_region_voting_chained_equals_oracle(cases):
    """All five voting calls back to back on the device (no re-seeding from the oracle)."""
    for name, l, r, d, st in cases:
        run, N = _runner(l, r, d)
        run.run(N.STAGE_PREP)
        run.set_disp(st.lrc)
        run.run(N.STAGE_VOTE, -1)
        assert np.array_equal(run.disp(), st.vote[4]), name
        run.close()


def _final_stats(got, want):
    diff = np.abs(got.astype(np.float64) - want.astype(np.float64))
    return dict(max=float(diff.max()), gt005=float((diff > 0.05).mean()), gt1=float((diff > 1.0).mean()))


def test_end_to_end_within_tolerance(cases):
    import tea_stereo_matching_b200 as t

    for name, l, r, d, st in cases:
        m = t.ADCensus()
        m.setMatchingStrategy(t.ColorModel.RGB, False, False)
        m.setMinMaxDisparity(0, d)
        got = m.compute(l, r)
        s = _final_stats(got, st.final)
        assert s["gt1"] <= 1e-3, (name, s)
        assert s["gt005"] <= 1e-3, (name, s)
        # deterministic: a second call returns the same bits
        assert np.array_equal(got, m.compute(l, r)), name


def test_end_to_end_against_reference_golden(pair_0600, golden_0600, native_lib):
    """Same check against the committed output of the unmodified reference (no oracle port involved)."""
    import tea_stereo_matching_b200 as t

    left, right = pair_0600
    m = t.ADCensus()
    m.setMatchingStrategy(t.ColorModel.RGB)
    m.setMinMaxDisparity(0, int(golden_0600["max_disparity"]))
    got = m.compute(left, right)
    s = _final_stats(got, golden_0600["final"])
    assert s["gt1"] <= 1e-3 and s["gt005"] <= 1e-3, s


def test_unsupported_modes_are_explicit(pair_0600, native_lib):
    import tea_stereo_matching_b200 as t
    from tea_stereo_matching_b200 import _native as N

    left, right = pair_0600
    m = t.ADCensus()
    wide = np.zeros((8, 1600, 3), np.uint8)  # ROI / mask modes search W / 2 + 1 = 801 > 768 levels
    m.setMatchingStrategy(t.ColorModel.RGB, True, False)
    with pytest.raises(t.ADCensusError) as e:
        m.compute(wide, wide)
    assert e.value.status == N.TSM_E_UNSUPPORTED
    m.setMatchingStrategy(t.ColorModel.RGB, False, False)
    with pytest.raises(t.ADCensusError):  # the setter refuses min * max < 0 like the reference (ADCensus.cpp:309)
        m.setMinMaxDisparity(-4, 48)


def test_strided_inputs_and_async_pair(pair_0600, port, native_lib):
    import tea_stereo_matching_b200 as t

    left, right = pair_0600
    wide = np.zeros((left.shape[0], left.shape[1] + 13, 3), np.uint8)
    wide[:, : left.shape[1]] = left
    lv = wide[:, : left.shape[1]]  # non-contiguous rows (step > 3*W)
    m = t.ADCensus()
    m.setMatchingStrategy(t.ColorModel.RGB)
    m.setMinMaxDisparity(0, 32)
    a = m.compute(lv, right)
    m.enqueue(left, right)
    b = m.wait()
    assert np.array_equal(a, b)


def test_rectify_matches_cv_restatement(port, native_lib):
    """tsm_remap / tsm_rectify_stereo bit-exact vs the cv2-pinned restatement (oracle/cvport.c)."""
    import ctypes as C

    import tea_stereo_matching_b200 as t
    from tea_stereo_matching_b200.synth import convert_maps_fixed, synth_rectify_maps, synth_v1

    H, W = 96, 160
    l, r = synth_v1(H, W, 24, seed=21)
    stereo = np.ascontiguousarray(np.concatenate([l, r], axis=1))
    (mx0, my0), (mx1, my1) = synth_rectify_maps(H, W)
    # push part of the map outside the image to exercise BORDER_CONSTANT
    mx0 = mx0 - 9.3
    my1 = my1 + 7.7
    fixed = [convert_maps_fixed(mx0, my0), convert_maps_fixed(mx1, my1)]

    def cpu_remap(src, m1, m2):
        out = np.empty((H, W, 3), np.uint8)
        src = np.ascontiguousarray(src)
        port.lib.cvp_remap_bilinear_8uc3_fixed(src.ctypes.data_as(C.c_void_p), src.shape[0], src.shape[1], src.strides[0],
                                               m1.ctypes.data_as(C.c_void_p), m2.ctypes.data_as(C.c_void_p),
                                               out.ctypes.data_as(C.c_void_p), H, W)
        return out

    want = [cpu_remap(l, *fixed[0]), cpu_remap(r, *fixed[1])]
    for maps in (
        t.EpipolarRectifyMap(map00=fixed[0][0], map01=fixed[0][1], map10=fixed[1][0], map11=fixed[1][1]),
        t.EpipolarRectifyMap(map00=mx0, map01=my0, map10=mx1, map11=my1),
    ):
        rect = t.EpipolarRectify(maps, (W, H))
        gl, gr = rect.rectify(stereo)
        assert np.array_equal(gl, want[0]) and np.array_equal(gr, want[1])
        gl2, gr2 = rect.rectify(l, r)
        assert np.array_equal(gl2, want[0]) and np.array_equal(gr2, want[1])
        both = rect.rectifyStereo(stereo)
        assert np.array_equal(both, np.concatenate(want, axis=1))
    # fused rectify -> ADCensus == ADCensus on the separately rectified pair
    m = t.ADCensus()
    m.setMatchingStrategy(t.ColorModel.RGB)
    m.setMinMaxDisparity(0, 24)
    fused = rect.rectify_adcensus(stereo, m)
    assert np.array_equal(fused, m.compute(want[0], want[1]))


def test_division_shortcut_is_ieee_exact(native_lib):
    """div_exact (branch-free fast path) == __fdiv_rn for every divisor the normalisation can see."""
    import ctypes as C

    import tea_stereo_matching_b200 as t

    ctx = t.Context(0)
    bad = C.c_ulonglong(123)
    ctx.check(native_lib.tsm_selftest(ctx.handle, 0, C.byref(bad)))
    assert bad.value == 0
    ctx.close()
