"""Run-to-run determinism with several contexts in flight on one GPU.

Regression test for a cross-proxy hazard in the scanline's TMA pipeline: with a shared-memory-heavy kernel of
another stream (the aggregation walk) co-resident on the SM, a stage refill overtook the warp's outstanding
ld.shared and ~95 % of the runs differed in a few hundred pixels.  `scripts/stress_stages.py` is the long form.
"""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_pipeline_is_deterministic_next_to_other_contexts(pair_0600):
    import tea_stereo_matching_b200 as t
    from tea_stereo_matching_b200 import _native as N
    from tea_stereo_matching_b200.adcensus import StageRunner

    left, right = pair_0600
    stop = threading.Event()

    def noise(D):
        run = StageRunner(left, right, D)
        run.run(N.STAGE_PREP | N.STAGE_INIT)
        while not stop.is_set():
            run.run(N.STAGE_AGGREGATE)
        run.close()

    # D = 95 / 48: the geometry the hazard was first seen with; D = 192: the benchmark's disparity range on both sides
    ths = [threading.Thread(target=noise, args=(D,)) for D in (95, 48, 192)]
    [th.start() for th in ths]
    try:
        for D in (48, 31, 64, 192):
            m = t.ADCensus()
            m.setMatchingStrategy(t.ColorModel.RGB)
            m.setMinMaxDisparity(0, D)
            ref = m.compute(left, right)
            for it in range(25):
                got = m.compute(left, right)
                assert np.array_equal(got, ref), f"D={D} run {it}: {(got != ref).sum()} pixels differ"
    finally:
        stop.set()
        [th.join() for th in ths]


def test_batched_enqueue_matches_single(pair_0600):
    import tea_stereo_matching_b200 as t

    left, right = pair_0600
    ms = []
    for _ in range(3):
        m = t.ADCensus()
        m.setMatchingStrategy(t.ColorModel.RGB)
        m.setMinMaxDisparity(0, 48)
        ms.append(m)
    ref = ms[0].compute(left, right)
    for it in range(10):
        for m in ms:
            m.enqueue(left, right)
        for k, m in enumerate(ms):
            assert np.array_equal(m.wait(), ref), f"iteration {it}, context {k}"


def test_context_reuse_across_geometries_and_models(pair_0600):
    """One context, arena grown once: a big RGB pair, then smaller pairs / other models must equal fresh contexts."""
    import tea_stereo_matching_b200 as t
    from tea_stereo_matching_b200.synth import synth_v1

    left, right = pair_0600
    big = synth_v1(200, 700, 192, seed=41)
    jobs = [(big[0], big[1], 192, t.ColorModel.RGB), (left, right, 48, t.ColorModel.RGB), (left, right, 48, t.ColorModel.HSI),
            (left[10:120, 30:250].copy(), right[10:120, 30:250].copy(), 70, t.ColorModel.RGB), (left, right, 31, t.ColorModel.HSI),
            (big[0], big[1], 128, t.ColorModel.RGB)]
    shared = t.ADCensus()
    for l, r, D, model in jobs:
        shared.setMatchingStrategy(model)
        shared.setMinMaxDisparity(0, D)
        got = shared.compute(l, r)
        fresh = t.ADCensus()
        fresh.setMatchingStrategy(model)
        fresh.setMinMaxDisparity(0, D)
        assert np.array_equal(got, fresh.compute(l, r)), (l.shape, D, model)


def test_two_devices_in_one_process(pair_0600):
    """One process driving two GPUs: function attributes (dynamic shared memory limits) are per device."""
    import tea_stereo_matching_b200 as t
    from tea_stereo_matching_b200 import _native as N

    import ctypes as C

    n = C.c_int32(0)
    N.lib().tsm_device_count(C.byref(n))
    if n.value < 2:
        pytest.skip("needs two GPUs")
    left, right = pair_0600
    outs = []
    for dev in (0, 1):
        for model in (t.ColorModel.RGB, t.ColorModel.HSI):
            m = t.ADCensus(device=dev)
            m.setMatchingStrategy(model)
            m.setMinMaxDisparity(0, 64)
            outs.append(m.compute(left, right))
    assert np.array_equal(outs[0], outs[2]) and np.array_equal(outs[1], outs[3])


def test_batch_sharded_over_devices_equals_one_device(pair_0600):
    """every frame of a batch sharded over N GPUs equals the 1-GPU bits (thread per device, four
    pairs in flight each).  With one visible device the same code path runs with N = 1 worker."""
    import ctypes as C

    import tea_stereo_matching_b200 as t
    from tea_stereo_matching_b200 import _native as N
    from tea_stereo_matching_b200.synth import synth_v1

    left, right = pair_0600
    pairs = [(left, right)] + [synth_v1(180, 320, 48, seed=70 + i) for i in range(6)]
    m = t.ADCensus(device=0)
    m.setMatchingStrategy(t.ColorModel.RGB)
    m.setMinMaxDisparity(0, 48)
    single = [m.compute(l, r) for l, r in pairs]
    one = m.computeBatch([p[0] for p in pairs], [p[1] for p in pairs])
    assert all(np.array_equal(a, b) for a, b in zip(single, one))
    n = C.c_int32(0)
    N.lib().tsm_device_count(C.byref(n))
    allgpu = m.computeBatch([p[0] for p in pairs], [p[1] for p in pairs], devices=-1)
    assert all(np.array_equal(a, b) for a, b in zip(single, allgpu)), f"{n.value} devices"
    # an error in the middle of a batch leaves the contexts usable (no un-waited pair)
    bad = [p[0] for p in pairs]
    bad[3] = bad[3][:, :-1]
    with pytest.raises(t.ADCensusError):
        m.computeBatch(bad, [p[1] for p in pairs])
    again = m.computeBatch([p[0] for p in pairs], [p[1] for p in pairs])
    assert all(np.array_equal(a, b) for a, b in zip(single, again))
