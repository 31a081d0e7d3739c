#!/usr/bin/env python
"""CPU emulation of the aggregation arithmetic of k_agg_fused (csrc/k_aggregate.cu): 48-bit wrap-around
fixed-point prefix sums (scale 2^34 in a normalising pass, 2^40 otherwise) instead of fp64 prefixes, checked against
the oracle's sequential fp32 sums (reference source/ADCensus.cpp:685-751).  Run in the dev container (no GPU):

    python scripts/emulate_agg_fixed.py

Prints the maximum relative difference per test pair for (a) fp64 prefixes everywhere (what k_agg_persist does) and
(b) fixed-point prefixes in the three fused launches (V.norm+V, H.norm+H, V.norm+V).  Bar: 2e-6 (tests/test_gpu_parity.py)."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import oracle  # noqa: E402
from tea_stereo_matching_b200.synth import synth_v1  # noqa: E402

MASK48 = np.uint64((1 << 48) - 1)


def window_counts(arms):
    up, down, left, right = [a.astype(np.int64) for a in arms]
    H, W = up.shape
    hlen = left + right + 1
    vlen = up + down + 1
    # N after (H then V): sum over the vertical arm of hlen ; after (V then H): sum over the horizontal arm of vlen
    ch = np.concatenate([np.zeros((1, W), np.int64), np.cumsum(hlen, axis=0)], axis=0)
    cv = np.concatenate([np.zeros((H, 1), np.int64), np.cumsum(vlen, axis=1)], axis=1)
    ys, xs = np.mgrid[0:H, 0:W]
    n_hv = ch[ys + down + 1, xs] - ch[ys - up, xs]
    n_vh = cv[ys, xs + right + 1] - cv[ys, xs - left]
    return n_hv, n_vh


def one_pass(vol, a, b, axis, norm, fixed):
    """out(p) = sum_{j=-a..b} in(p + j) along `axis` (0 vertical, 1 horizontal) as a prefix difference rounded once to fp32,
    then / N (fp32 IEEE) when norm is given."""
    v = np.moveaxis(vol, axis, 0)  # [len][other][D]
    L = v.shape[0]
    am = np.moveaxis(a, axis, 0)[:, :, None]
    bm = np.moveaxis(b, axis, 0)[:, :, None]
    idx = np.arange(L)[:, None, None]
    if fixed:
        S = 34 if norm is not None else 40
        fx = np.rint(v.astype(np.float64) * float(1 << S)).astype(np.int64).astype(np.uint64)
        P = np.concatenate([np.zeros((1,) + v.shape[1:], np.uint64), np.cumsum(fx, axis=0, dtype=np.uint64)], axis=0)
        hi = np.take_along_axis(P, np.broadcast_to(idx + bm + 1, v.shape), axis=0)
        lo = np.take_along_axis(P, np.broadcast_to(idx - am, v.shape), axis=0)
        d = (hi - lo) & MASK48
        out = (d.astype(np.float64) * (1.0 / float(1 << S))).astype(np.float32)
    else:
        P = np.concatenate([np.zeros((1,) + v.shape[1:], np.float64), np.cumsum(v.astype(np.float64), axis=0)], axis=0)
        hi = np.take_along_axis(P, np.broadcast_to(idx + bm + 1, v.shape), axis=0)
        lo = np.take_along_axis(P, np.broadcast_to(idx - am, v.shape), axis=0)
        out = (hi - lo).astype(np.float32)
    if norm is not None:
        out = out / np.moveaxis(norm, axis, 0)[:, :, None].astype(np.float32)
    return np.ascontiguousarray(np.moveaxis(out, 0, axis))


def aggregate(vol, arms, fused_fixed):
    up, down, left, right = [x.astype(np.int64) for x in arms]
    n_hv, n_vh = window_counts(arms)
    # pass list: (axis, norm); launches: H | Vn+V | Hn+H | Vn+V | Hn
    passes = [(1, None), (0, n_hv), (0, None), (1, n_vh), (1, None), (0, n_hv), (0, None), (1, n_vh)]
    fused = [False, True, True, True, True, True, True, False]
    for (axis, norm), fz in zip(passes, fused):
        a, b = (up, down) if axis == 0 else (left, right)
        vol = one_pass(vol, a, b, axis, norm, fixed=fused_fixed and fz)
    return vol


def main():
    port = oracle.Port()
    z = np.load(ROOT / "tests" / "golden" / "pair_0600_320x180.npz")
    cases = [("0600_320x180_d48", z["left"], z["right"], 48),
             ("synth_72x420_d64",) + synth_v1(72, 420, 64, seed=32) + (64,),
             ("synth_gray_120x300_d40",) + synth_v1(120, 300, 40, seed=34, gray=True) + (40,)]
    # saturated regions: exact zero costs next to tiny averages
    l, r = synth_v1(90, 260, 32, seed=5)
    l = l.copy(); r = r.copy()
    l[20:70, 40:200] = 255; r[20:70, 30:190] = 255
    l[:, 230:] = 0; r[:, 225:] = 0
    cases.append(("synth_saturated_90x260_d32", l, r, 32))
    worst = 0.0
    for name, left, right, D in cases:
        st = port.run(left, right, D)
        for view in range(2):
            want = st.vol_agg[view]
            for label, fz in (("fp64", False), ("fixed48", True)):
                got = aggregate(st.vol_init[view], st.arms[view], fz)
                rel = np.abs(got - want) / np.maximum(np.abs(want), 1e-30)
                zero_mismatch = int(((got == 0) != (want == 0)).sum())
                print(f"{name} view {view} {label:8s} max rel {rel.max():.3e}  min nonzero want {want[want > 0].min():.3e}  "
                      f"zero/nonzero mismatches {zero_mismatch}", flush=True)
                if fz:
                    worst = max(worst, float(rel.max()))
    print("worst fixed48:", worst, "bar 2e-6:", "OK" if worst <= 2e-6 else "FAIL")


if __name__ == "__main__":
    main()
