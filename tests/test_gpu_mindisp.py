"""setMinMaxDisparity(min > 0, max) -- the reference's handling is odd but defined (plane index vs disparity mix,
ADCensus.cpp:556-561, :890, :1398, :1310-1322, :1358): the CUDA path reproduces it as it is, stage by stage.

* committed golden of the UNMODIFIED reference, (4, 28) on a 128 x 48 crop with full volumes
  (tests/golden/make_mindisp_golden.py) -- runs on a box without /root/reference;
* live against oracle/_ref (the reference compiled here; travels to the GPU box) for (4, 48) and (16, 128)."""
import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

AGG_RTOL = 2e-6


def _check_stages(left, right, mind, maxd, g):
    """g: mapping with vol_init{0,1}, vol_agg{0,1}, vol_scan{0,1}, arms{0,1}, wta{0,1}, lrc, vote0..4, interp, discont, final."""
    import tea_stereo_matching_b200 as t
    from tea_stereo_matching_b200 import _native as N

    run = t.StageRunner(left, right, maxd, min_disparity=mind)
    run.run(N.STAGE_PREP | N.STAGE_INIT)
    for v in range(2):
        assert np.array_equal(run.arms(v), g[f"arms{v}"]), ("arms", v)
        got = run.volume(v)
        assert np.array_equal(got, g[f"vol_init{v}"]), ("vol_init", v, int((got != g[f"vol_init{v}"]).sum()))
    run.run(N.STAGE_AGGREGATE)
    for v in range(2):
        got, want = run.volume(v), g[f"vol_agg{v}"]
        rel = np.abs(got - want) / np.maximum(np.abs(want), 1e-30)
        assert rel.max() <= AGG_RTOL, ("vol_agg", v, float(rel.max()))
    for v in range(2):
        run.set_volume(v, g[f"vol_agg{v}"])
    run.run(N.STAGE_SCANLINE)
    for v in range(2):
        got = run.volume(v)
        assert np.array_equal(got, g[f"vol_scan{v}"]), ("vol_scan", v, int((got != g[f"vol_scan{v}"]).sum()))
        assert np.array_equal(run.wta(v), g[f"wta{v}"]), ("fused wta", v)
    run.run(N.STAGE_WTA)  # the stand-alone kernel
    for v in range(2):
        assert np.array_equal(run.wta(v), g[f"wta{v}"]), ("wta", v)
    run.run(N.STAGE_LRC)
    assert np.array_equal(run.disp(), g["lrc"]), "lrc"
    prev = g["lrc"]
    for i in range(5):
        run.set_disp(prev)
        run.run(N.STAGE_VOTE, i)
        got = run.disp()
        assert np.array_equal(got, g[f"vote{i}"]), ("vote", i, int((got != g[f"vote{i}"]).sum()))
        prev = g[f"vote{i}"]
    run.set_disp(g["vote4"])
    run.run(N.STAGE_INTERP)
    assert np.array_equal(run.disp(), g["interp"]), "interp"
    run.set_disp(g["interp"])
    run.run(N.STAGE_DISCONT)
    assert np.array_equal(run.disp(), g["discont"]), "discont"
    run.set_disp(g["discont"])
    run.run(N.STAGE_SUBPIXEL)
    assert np.array_equal(run.final(), g["final"]), "final"
    run.close()
    # the public operator, end to end
    m = t.ADCensus()
    m.setMatchingStrategy(t.ColorModel.RGB, False, False)
    m.setMinMaxDisparity(mind, maxd)
    got = m.compute(left, right)
    diff = np.abs(got.astype(np.float64) - g["final"].astype(np.float64))
    assert (diff > 1.0).mean() <= 1e-3 and (diff > 0.05).mean() <= 1e-3, (float(diff.max()), float((diff > 0.05).mean()))


def test_min_disparity_vs_committed_reference_golden(native_lib):
    z = np.load(GOLDEN / "ref_0600_crop_128x48_d4_28.npz")
    g = {k: (z[k].astype(np.int32) if z[k].dtype == np.int16 else z[k]) for k in z.files}
    _check_stages(z["left"], z["right"], int(z["min_disparity"]), int(z["max_disparity"]), g)


@pytest.mark.parametrize("mind,maxd,which", [(4, 48, "0600"), (16, 128, "synth")])
def test_min_disparity_vs_reference_live(mind, maxd, which, ref, pair_0600, native_lib):
    from tea_stereo_matching_b200.synth import synth_v1

    left, right = pair_0600 if which == "0600" else synth_v1(72, 420, maxd, seed=32)
    st = ref.run(left, right, maxd, serial_scanline=True, min_disp=mind)
    g = dict(lrc=st.lrc, interp=st.interp, discont=st.discont, final=st.final)
    for v in range(2):
        g[f"vol_init{v}"], g[f"vol_agg{v}"], g[f"vol_scan{v}"] = st.vol_init[v], st.vol_agg[v], st.vol_scan[v]
        g[f"arms{v}"] = np.stack(st.arms[v], axis=2).astype(np.uint8)
        g[f"wta{v}"] = st.wta[v]
    for i in range(5):
        g[f"vote{i}"] = st.vote[i]
    _check_stages(left, right, mind, maxd, g)


def test_empty_wta_range_is_refused(pair_0600, native_lib):
    """max < 2 min: the reference's cost2disparity loop (minD .. maxD - minD) is empty and it returns uninitialised memory."""
    import tea_stereo_matching_b200 as t
    from tea_stereo_matching_b200 import _native as N

    left, right = pair_0600
    m = t.ADCensus()
    m.setMatchingStrategy(t.ColorModel.RGB, False, False)
    m.setMinMaxDisparity(30, 48)
    with pytest.raises(t.ADCensusError) as e:
        m.compute(left, right)
    assert e.value.status == N.TSM_E_UNSUPPORTED
