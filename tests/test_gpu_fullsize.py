"""Full-size parity of all five BASELINE configs at their real geometry, plus size-independent properties.

C1 / C2 : committed outputs of the UNMODIFIED reference (tests/golden/ref_c*_d*.npz, serial-scanline semantics,
          made by make_golden.py full); C1 also stage by stage with the oracle port on the GPU box's host cores.
C3      : frame seed = 1000 of the benchmark batch (1080p, D = 0..192) against oracle.Port().compute run live.
C4      : fused rectify -> ADCensus on the 2 x 1280x1024 side-by-side frame (D = 0..128) against the cv2-pinned
          remap restatement (oracle/cvport.c) followed by the port, live.
C5      : 4K gray, D = 0..384, against the committed output of the port (tests/golden/port_c5_gray4k_d384.npz,
          made by make_c5_golden.py: every 4th row of the final and WTA maps + the full invalid-pixel mask).
Bars (north_star): at most 0.1 % of pixels off by more than 1 px, at most 0.1 % off by more than 0.05 px.
Every measured (max, >0.05 px, >1 px) triple is appended to gpurun_out/parity_r2.json (copied to profiles/)."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

pytestmark = pytest.mark.gpu

BAR_GT1 = 1e-3
BAR_GT005 = 1e-3


def _stats(got, want):
    diff = np.abs(got.astype(np.float64) - want.astype(np.float64))
    return dict(max=float(diff.max()), gt005=float((diff > 0.05).mean()), gt1=float((diff > 1.0).mean()),
                invalid_mismatch=float(((got < 0) != (want < 0)).mean()))


def _record(name, stats, **extra):
    """Appends the measured parity of one config to gpurun_out/parity_r2.json (best effort)."""
    out = ROOT / "gpurun_out"
    try:
        out.mkdir(exist_ok=True)
        f = out / "parity_r2.json"
        data = json.loads(f.read_text()) if f.exists() else {}
        data[name] = dict(stats, bars={"gt1": BAR_GT1, "gt005": BAR_GT005}, **extra)
        f.write_text(json.dumps(data, indent=1, sort_keys=True))
    except OSError:
        pass


def _matcher(maxd):
    import tea_stereo_matching_b200 as t

    m = t.ADCensus()
    m.setMatchingStrategy(t.ColorModel.RGB, False, False)
    m.setMinMaxDisparity(0, maxd)
    return m


@pytest.mark.parametrize("name,maxd", [("c1_0600_720p", 192), ("c2_motorcycle", 256)])
def test_full_size_vs_reference_golden(name, maxd, native_lib):
    pair = np.load(GOLDEN / f"pair_{name}.npz")
    gold = np.load(GOLDEN / f"ref_{name}_d{maxd}.npz")
    got = _matcher(maxd).compute(pair["left"], pair["right"])
    s = _stats(got, gold["final"])
    print(name, s)
    _record(name, s, against="unmodified reference (committed golden)", shape=list(got.shape), max_disparity=maxd)
    # north_star: within 0.05 px, at most 0.1 % of pixels differing by more than 1 px
    assert s["gt1"] <= BAR_GT1, s
    assert s["gt005"] <= BAR_GT005, s


def test_c1_stage_by_stage_vs_oracle_port(port, native_lib):
    """720p, D = 0..192: every stage on the oracle's input for that stage."""
    import tea_stereo_matching_b200 as t
    from tea_stereo_matching_b200 import _native as N

    pair = np.load(GOLDEN / "pair_c1_0600_720p.npz")
    left, right = pair["left"], pair["right"]
    st = port.run(left, right, 192)
    gold = np.load(GOLDEN / "ref_c1_0600_720p_d192.npz")
    assert np.array_equal(st.final, gold["final"])  # the port still equals the reference at full size
    run = t.StageRunner(left, right, 192)
    run.run(N.STAGE_PREP | N.STAGE_INIT)
    for v in range(2):
        assert np.array_equal(run.arms(v), np.stack(st.arms[v], axis=2))
        assert np.array_equal(run.volume(v), st.vol_init[v])
    run.run(N.STAGE_AGGREGATE)
    for v in range(2):
        got, want = run.volume(v), st.vol_agg[v]
        rel = np.abs(got - want) / np.maximum(np.abs(want), 1e-30)
        print("agg rel err view", v, float(rel.max()))
        assert rel.max() <= 2e-6
    for v in range(2):
        run.set_volume(v, st.vol_agg[v])
    run.run(N.STAGE_SCANLINE)
    for v in range(2):
        assert np.array_equal(run.volume(v), st.vol_scan[v])
        assert np.array_equal(run.wta(v), st.wta[v])
    run.run(N.STAGE_LRC)
    assert np.array_equal(run.disp(), st.lrc)
    run.run(N.STAGE_VOTE, -1)
    assert np.array_equal(run.disp(), st.vote[4])
    run.run(N.STAGE_INTERP)
    assert np.array_equal(run.disp(), st.interp)
    run.run(N.STAGE_DISCONT)
    assert np.array_equal(run.disp(), st.discont)
    run.run(N.STAGE_SUBPIXEL)
    assert np.array_equal(run.final(), st.final)
    run.close()


def test_c3_frame_vs_oracle_port_live(port, native_lib):
    """The headline config: frame 0 of the benchmark batch (seed 1000), 1080p, D = 0..192, against the port."""
    from tea_stereo_matching_b200.synth import synth_v1

    H, W, D = 1080, 1920, 192
    left, right = synth_v1(H, W, D, seed=1000)
    got = _matcher(D).compute(left, right)
    want = port.compute(left, right, D)
    s = _stats(got, want)
    print("c3", s)
    _record("c3_synth_1080p_seed1000", s, against="oracle port, live", shape=[H, W], max_disparity=D)
    assert s["gt1"] <= BAR_GT1 and s["gt005"] <= BAR_GT005, s


def test_c4_fused_rectify_adcensus_vs_cv_remap_plus_port_live(port, native_lib):
    """Config C4 at its real size: hconcat(synth_v1(1024, 1280, 128, seed 2000)) + the SURVEY 8(d) synthetic maps."""
    import ctypes as C

    import tea_stereo_matching_b200 as t
    from tea_stereo_matching_b200.synth import convert_maps_fixed, synth_rectify_maps, synth_v1

    H, W, D = 1024, 1280, 128
    l, r = synth_v1(H, W, D, seed=2000)
    stereo = np.ascontiguousarray(np.concatenate([l, r], axis=1))
    (mx0, my0), (mx1, my1) = synth_rectify_maps(H, W)
    fixed = [convert_maps_fixed(mx0, my0), convert_maps_fixed(mx1, my1)]

    def cpu_remap(src, m1, m2):
        out = np.empty((H, W, 3), np.uint8)
        src = np.ascontiguousarray(src)
        port.lib.cvp_remap_bilinear_8uc3_fixed(src.ctypes.data_as(C.c_void_p), src.shape[0], src.shape[1], src.strides[0],
                                               m1.ctypes.data_as(C.c_void_p), m2.ctypes.data_as(C.c_void_p),
                                               out.ctypes.data_as(C.c_void_p), H, W)
        return out

    rl, rr = cpu_remap(l, *fixed[0]), cpu_remap(r, *fixed[1])
    maps = t.EpipolarRectifyMap(map00=fixed[0][0], map01=fixed[0][1], map10=fixed[1][0], map11=fixed[1][1])
    rect = t.EpipolarRectify(maps, (W, H))
    gl, gr = rect.rectify(stereo)
    assert np.array_equal(gl, rl) and np.array_equal(gr, rr)  # remap is integer fixed point: bit-exact at full size
    m = _matcher(D)
    got = rect.rectify_adcensus(stereo, m)
    want = port.compute(rl, rr, D)
    s = _stats(got, want)
    print("c4", s)
    _record("c4_rectify_1280x1024_seed2000", s, against="cv2-pinned remap restatement + oracle port, live", shape=[H, W],
            max_disparity=D, remap_bit_exact=True)
    assert s["gt1"] <= BAR_GT1 and s["gt005"] <= BAR_GT005, s
    # a reloaded map at the same host addresses must not reuse the stale device copy (map_generation, include/tsm.h)
    fixed[0][0][...] = np.roll(fixed[0][0], 3, axis=1)
    rect.loadEpipolarRectifyMap(maps, (W, H))
    gl2, _ = rect.rectify(stereo)
    assert np.array_equal(gl2, cpu_remap(l, *fixed[0]))


def test_c3_geometry_properties(native_lib):
    """1080p, D = 0..192: determinism, value range, and a fronto-parallel plane is recovered."""
    from tea_stereo_matching_b200.synth import synth_v1

    H, W, D = 1080, 1920, 192
    left, right = synth_v1(H, W, D, seed=1000)
    m = _matcher(D)
    a = m.compute(left, right)
    b = m.compute(left, right)
    assert np.array_equal(a, b)
    assert np.isfinite(a).all() and a.max() <= D and a.min() >= -2.0
    assert (a >= 0).mean() > 0.9
    # constant-shift pair: right(x) = left(x + d0)  =>  disparity d0 wherever both views see the pixel
    d0 = 37
    shifted = np.zeros_like(left)
    shifted[:, : W - d0] = left[:, d0:]
    c = m.compute(left, shifted)
    inner = c[40:-40, d0 + 60 : W - 60]
    assert np.median(inner) == pytest.approx(d0, abs=0.05)
    assert (np.abs(inner - d0) <= 1.0).mean() > 0.97


def test_c5_vs_port_golden(native_lib):
    """4K gray, D = 0..384 (25.5 GB of cost volume, > 2^31 cells per view: the 64-bit indexing path) against the
    committed output of the oracle port."""
    import hashlib

    import tea_stereo_matching_b200 as t
    from tea_stereo_matching_b200 import _native as N
    from tea_stereo_matching_b200.synth import synth_v1

    g = np.load(GOLDEN / "port_c5_gray4k_d384.npz")
    H, W, D, step = int(g["H"]), int(g["W"]), int(g["max_disparity"]), int(g["row_step"])
    left, right = synth_v1(H, W, D, seed=int(g["seed"]), gray=True)
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    assert sha(left) == str(g["left_sha256"]) and sha(right) == str(g["right_sha256"]), "synth_v1 changed: regenerate the golden"
    m = _matcher(D)
    a = m.compute(left, right)
    assert np.isfinite(a).all() and a.max() <= D and a.min() >= -2.0
    s = _stats(a[::step], g["final_rows"])
    invalid = np.unpackbits(g["invalid_bits"])[: H * W].reshape(H, W).astype(bool)
    s["invalid_mismatch"] = float(((a < 0) != invalid).mean())  # full resolution
    s["bit_identical_to_port"] = sha(a) == str(g["final_sha256"])
    print("c5", s)
    _record("c5_gray4k_seed3000", s, against="oracle port (committed golden: every 4th row + full invalid mask)", shape=[H, W],
            max_disparity=D)
    assert s["gt1"] <= BAR_GT1 and s["gt005"] <= BAR_GT005, s
    assert s["invalid_mismatch"] <= BAR_GT1, s
    assert np.array_equal(a, m.compute(left, right))
