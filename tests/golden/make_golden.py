#!/usr/bin/env python
"""Generates the committed golden vectors under tests/golden/.

Run in the dev container (needs /root/reference and python cv2 4.13.0):
    python tests/golden/make_golden.py          # small fixtures (seconds)
    python tests/golden/make_golden.py full     # full-size C1 / C2 reference outputs (minutes)

* pair_0600_320x180.npz  -- the reference's demo pair demo-imgs/0600-{Left,Right}.bmp,
  area-downscaled x4 with cv2 (inputs only; BGR uint8).
* ref_0600_320x180_d48.npz -- every stage output of the UNMODIFIED reference
  source/ADCensus.cpp (oracle/_ref/libadcensus_ref.so, serial-scanline semantics) on that
  pair, D = 0..48: integer maps in full, float volumes as SHA-256 digests plus strided
  samples (the full volumes are 11 MB each).
* ref_0600_crop_160x96_d32_hsi.npz / ref_0600_320x180_d48_hsi.npz -- the same with the HSI colour model
  (bgr2hsi + computeGaussMedian preprocessing; a 160x96 crop carries the full volumes).
* ref_0600_crop_160x96_roi.npz -- final maps of the public compute() in ROI matching mode (RGB + offset 5, HSI + offset 3).
* ref_synth_96x128_d24.npz -- same for a tiny synthetic pair (synth_v1 seed 7), with the
  full volumes (small enough) so the CUDA kernels can be checked cell by cell on a box
  without /root/reference.
* cv_golden.npz -- cv2 4.13.0 outputs of equalizeHist / blur / Canny / medianBlur / remap /
  convertMaps on fixed random inputs, for oracle/cvport.c.
"""
import hashlib
import sys
from pathlib import Path

import cv2
import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import oracle  # noqa: E402
from tea_stereo_matching_b200.synth import synth_v1  # noqa: E402

OUT = Path(__file__).resolve().parent
REF_IMGS = Path("/root/reference/demo-imgs")


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def stage_dict(st, full_volumes: bool):
    d = {}
    for name in ("vol_init", "vol_agg", "vol_scan"):
        for k in range(2):
            v = getattr(st, name)[k]
            d[f"{name}{k}_sha256"] = np.array(sha(v))
            d[f"{name}{k}_sample"] = v[::7, ::5, :].copy()
            if full_volumes:
                d[f"{name}{k}"] = v
    for k in range(2):
        d[f"arms{k}"] = np.stack(st.arms[k], axis=2).astype(np.uint8)
        d[f"wta{k}"] = st.wta[k].astype(np.int16)
    d["lrc"] = st.lrc.astype(np.int16)
    for i in range(5):
        d[f"vote{i}"] = st.vote[i].astype(np.int16)
    d["interp"] = st.interp.astype(np.int16)
    d["discont"] = st.discont.astype(np.int16)
    d["final"] = st.final
    return d


def main():
    assert cv2.__version__.startswith("4.13"), cv2.__version__
    oracle.build()
    ref = oracle.Ref()

    L = cv2.imread(str(REF_IMGS / "0600-Left.bmp"))
    R = cv2.imread(str(REF_IMGS / "0600-Right.bmp"))
    Ls = cv2.resize(L, (320, 180), interpolation=cv2.INTER_AREA)
    Rs = cv2.resize(R, (320, 180), interpolation=cv2.INTER_AREA)
    np.savez_compressed(OUT / "pair_0600_320x180.npz", left=Ls, right=Rs)
    st = ref.run(Ls, Rs, 48, serial_scanline=True)
    np.savez_compressed(OUT / "ref_0600_320x180_d48.npz", max_disparity=48, **stage_dict(st, False))

    # HSI colour model (row f1): the preprocessed images the stages see + every stage output, with full volumes on a
    # crop small enough to commit
    Lc, Rc = np.ascontiguousarray(Ls[40:136, 96:256]), np.ascontiguousarray(Rs[40:136, 96:256])
    st = ref.run(Lc, Rc, 32, serial_scanline=True, model="HSI")
    np.savez_compressed(OUT / "ref_0600_crop_160x96_d32_hsi.npz", left=Lc, right=Rc, max_disparity=32, pre0=st.pre[0], pre1=st.pre[1],
                        **stage_dict(st, True))
    st = ref.run(Ls, Rs, 48, serial_scanline=True, model="HSI")
    np.savez_compressed(OUT / "ref_0600_320x180_d48_hsi.npz", max_disparity=48, pre0=st.pre[0], pre1=st.pre[1], **stage_dict(st, False))

    # ROI matching mode (row f1): maxD = W / 2, offset, final -1 marking (black left pixels); public compute, one thread
    Lr, Rr = Lc.copy(), Rc.copy()
    Lr[20:30, 40:60] = 0
    np.savez_compressed(OUT / "ref_0600_crop_160x96_roi.npz", left=Lr, right=Rr,
                        rgb_off5=ref.compute_ex(Lr, Rr, 64, "RGB", roi=True, offset=5),
                        hsi_off3=ref.compute_ex(Lr, Rr, 64, "HSI", roi=True, offset=3))

    # mask matching mode: black pixels are holes in BOTH images (a masked foreground, as the mode is meant for)
    Lm, Rm = Lc.copy(), Rc.copy()
    Lm[:, :24] = 0; Lm[60:75, 70:100] = 0; Lm[5, 50] = 0
    Rm[:, :10] = 0; Rm[58:74, 60:88] = 0; Rm[40, 120] = 0
    np.savez_compressed(OUT / "ref_0600_crop_160x96_mask.npz", left=Lm, right=Rm,
                        rgb_off2=ref.compute_ex(Lm, Rm, 64, "RGB", mask=True, offset=2),
                        hsi_off0=ref.compute_ex(Lm, Rm, 64, "HSI", mask=True, offset=0),
                        rgb_roi_mask=ref.compute_ex(Lm, Rm, 64, "RGB", roi=True, mask=True, offset=1))

    sl, sr = synth_v1(96, 128, 24, seed=7)
    st = ref.run(sl, sr, 24, serial_scanline=True)
    np.savez_compressed(OUT / "ref_synth_96x128_d24.npz", left=sl, right=sr, max_disparity=24, **stage_dict(st, True))

    # ---- cv2 golden vectors ----
    rng = np.random.default_rng(20261018)
    g = {}
    H, W = 61, 83
    img_a = rng.integers(0, 256, (H, W), dtype=np.uint8)
    img_b = cv2.GaussianBlur(rng.integers(0, 256, (H, W)).astype(np.float32), (0, 0), 3).astype(np.uint8)
    img_c = np.zeros((H, W), np.uint8)
    img_c[H // 3:, W // 4:] = 140
    img_c[: H // 2, : W // 2] += rng.integers(0, 60, (H // 2, W // 2), dtype=np.uint8)
    for name, img in (("a", img_a), ("b", img_b), ("c", img_c)):
        g[f"u8_{name}"] = img
        g[f"eq_{name}"] = cv2.equalizeHist(img)
        g[f"blur_{name}"] = cv2.blur(img, (3, 3))
        g[f"canny_{name}"] = cv2.Canny(img, 30, 90, apertureSize=3)
    f = rng.normal(0, 20, (H, W)).astype(np.float32)
    f[rng.random((H, W)) < 0.2] = -1
    f[rng.random((H, W)) < 0.1] = -2
    g["f32"] = f
    g["median"] = cv2.medianBlur(f, 3)
    sH, sW = 47, 65
    src = rng.integers(0, 256, (sH, sW, 3), dtype=np.uint8)
    mx = rng.uniform(-3, sW + 3, (H, W)).astype(np.float32)
    my = rng.uniform(-3, sH + 3, (H, W)).astype(np.float32)
    m1, m2 = cv2.convertMaps(mx, my, cv2.CV_16SC2)
    g["remap_src"], g["remap_mx"], g["remap_my"], g["remap_m1"], g["remap_m2"] = src, mx, my, m1, m2
    g["remap_fixed"] = cv2.remap(src, m1, m2, cv2.INTER_LINEAR)
    g["remap_float"] = cv2.remap(src, mx, my, cv2.INTER_LINEAR)
    np.savez_compressed(OUT / "cv_golden.npz", **g)
    for p in sorted(OUT.glob("*.npz")):
        print(p.name, p.stat().st_size)


if __name__ == "__main__" and len(sys.argv) == 1:
    main()


def full_size():
    """Full-size reference outputs (slow: ~3 min for C1, ~6 min for C2 on 8 cores)."""
    ref = oracle.Ref()
    for name, (lf, rf), maxd in (("c1_0600_720p", ("0600-Left.bmp", "0600-Right.bmp"), 192),
                                 ("c2_motorcycle", ("Motorcycle_Left.png", "Motorcycle_Right.png"), 256)):
        L = cv2.imread(str(REF_IMGS / lf))
        R = cv2.imread(str(REF_IMGS / rf))
        np.savez_compressed(OUT / f"pair_{name}.npz", left=L, right=R)
        st = ref.run(L, R, maxd, serial_scanline=True, volumes=False)
        np.savez_compressed(OUT / f"ref_{name}_d{maxd}.npz", max_disparity=maxd, wta0=st.wta[0].astype(np.int16),
                            wta1=st.wta[1].astype(np.int16), lrc=st.lrc.astype(np.int16), vote4=st.vote[4].astype(np.int16),
                            interp=st.interp.astype(np.int16), discont=st.discont.astype(np.int16), final=st.final,
                            seconds=np.array([st.seconds[k] for k in ("init", "agg", "scan", "multi")]))
        print(name, st.seconds)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "full":
    full_size()
