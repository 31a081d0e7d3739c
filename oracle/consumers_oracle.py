"""TEST INFRASTRUCTURE -- CPU restatement (numpy, one IEEE fp32 operation per numpy call, reference order) of the
disparity consumers of the reference's source/stereo.cpp:75-202.  Only tests/ may import this.

The reference functions need OpenCV C++ (cv::Mat iterators, cv::gemm, cv::divide) and are therefore not compiled
here; the two OpenCV calls are pinned against cv2 4.13 in tests/test_consumers_oracle.py (live when cv2 imports,
plus committed vectors tests/golden/consumers_cv_golden.npz made by tests/golden/make_consumers_golden.py).
"""
import numpy as np

F = np.float32


def jet_colormap() -> np.ndarray:
    """stereo::JETColorMap, stereo.cpp:75-93 (1x256 BGR)."""
    t = np.zeros((256, 3), np.int64)
    for i in range(32):
        t[i] = (128 + 4 * i, 0, 0)
    t[32] = (255, 0, 0)
    for i in range(63):
        t[33 + i] = (255, 4 + 4 * i, 0)
    t[96] = (254, 255, 2)
    for i in range(62):
        t[97 + i] = (250 - 4 * i, 255, 6 + 4 * i)
    t[159] = (1, 255, 254)
    for i in range(64):
        t[160 + i] = (0, 252 - 4 * i, 255)
    for i in range(32):
        t[224 + i] = (0, 0, 252 - 4 * i)
    return t.astype(np.uint8).reshape(1, 256, 3)


def _to_uchar_x86(t: np.ndarray) -> np.ndarray:
    """static_cast<unsigned char>(float) as gcc/x86-64 does it: cvttss2si (INT_MIN when NaN / out of range), low byte."""
    bad = ~np.isfinite(t) | (t >= F(2147483648.0)) | (t < F(-2147483648.0))
    i = np.where(bad, np.int64(-2147483648), np.trunc(np.where(bad, 0, t)).astype(np.int64))
    return (i & 0xFF).astype(np.uint8)


def apply_colormap(src: np.ndarray, color_map: np.ndarray, min_val=None, max_val=None) -> np.ndarray:
    """stereo::applyColorMap, stereo.cpp:95-118 (auto range) and :120-137 (explicit range)."""
    src = np.asarray(src, F)
    cm = np.asarray(color_map, np.uint8).reshape(256, 3)
    with np.errstate(all="ignore"):
        if min_val is None:
            use = ~((src < 0) | np.isinf(src) | np.isnan(src))  # NaN never wins a std::min / std::max
            mn = src[use].min() if use.any() else F(np.inf)
            mx = src[use].max() if use.any() else F(-np.inf)
            mn, mx = F(abs(mn)) if mn == 0 else F(mn), F(mx)
            black = src < 0
        else:
            mn, mx = F(min_val), F(max_val)
            black = (src < mn) | (src > mx)
        t = ((src - mn) / F(mx - mn)) * F(255)
    out = cm[_to_uchar_x86(t.astype(F))]
    out[black] = 0
    return out


def reproject_to_depth(disp: np.ndarray, focal: float, baseline: float) -> np.ndarray:
    """stereo::reprojectToDepth, stereo.cpp:139-151."""
    disp = np.asarray(disp, F)
    fb = F(focal) * F(baseline)
    skip = (disp < 0) | np.isinf(disp)
    with np.errstate(all="ignore"):
        depth = (fb / disp).astype(F)
    depth[skip] = 0
    return depth


def reproject_to_3d(disp: np.ndarray, focal: float, baseline: float, cx: float, cy: float) -> np.ndarray:
    """stereo::reprojectTo3D(disparity, focalLength, baseline, cx, cy, XYZ), stereo.cpp:153-172."""
    disp = np.asarray(disp, F)
    H, W = disp.shape
    fb = F(focal) * F(baseline)
    u = np.arange(W, dtype=np.int64).astype(F)[None, :]
    v = np.arange(H, dtype=np.int64).astype(F)[:, None]
    skip = (disp < 0) | np.isinf(disp)
    with np.errstate(all="ignore"):
        Z = (fb / disp).astype(F)
        zf = (Z / F(focal)).astype(F)
        X = ((u - F(cx)).astype(F) * zf).astype(F)
        Y = ((v - F(cy)).astype(F) * zf).astype(F)
    xyz = np.stack([X, Y, Z], axis=-1)
    xyz[skip] = 0
    return xyz


def gemm32f(A: np.ndarray, B: np.ndarray) -> np.ndarray:
    """cv::gemm on CV_32F operands as OpenCV 4.13 computes a 4x4 by 4xN product: fp32 products and fp32 sums,
    k ascending, no fused multiply-add (pinned against cv2.gemm on 800 000 columns: 0 mismatching bit patterns;
    double accumulation or FMA mismatch on 40 % of them)."""
    A, B = np.asarray(A, F), np.asarray(B, F)
    s = np.zeros((A.shape[0], B.shape[1]), F)
    with np.errstate(all="ignore"):
        for k in range(A.shape[1]):
            s = (s + (A[:, k : k + 1] * B[k : k + 1, :]).astype(F)).astype(F)
    return s


def reproject_to_3d_q(disp: np.ndarray, Q: np.ndarray) -> np.ndarray:
    """stereo::reprojectTo3D(disparity, Q, XYZ), stereo.cpp:174-202."""
    disp = np.asarray(disp, F)
    H, W = disp.shape
    u = np.broadcast_to(np.arange(W, dtype=F)[None, :], (H, W)).reshape(1, -1)
    v = np.broadcast_to(np.arange(H, dtype=F)[:, None], (H, W)).reshape(1, -1)
    pix = np.concatenate([u, v, disp.reshape(1, -1), np.ones((1, H * W), F)], axis=0)
    Q32 = np.asarray(Q, np.float64).astype(F)
    xyzw = gemm32f(Q32, pix)
    with np.errstate(all="ignore"):
        xyz = (xyzw[:3] / xyzw[3:4]).astype(F)  # cv::divide on floats: plain IEEE division
    return np.ascontiguousarray(xyz.T.reshape(H, W, 3))


# ---- writePointCloudToPCD / writePointCloudToPLY (source/stereo.cpp:204-356) ------------------------------------
def to_chars_f32(x) -> str:
    """std::to_chars(first, last, float) (C++17 [charconv.to.chars]): the shortest decimal representation that round-trips,
    printed in fixed or scientific notation, whichever is shorter (fixed on a tie)."""
    x = np.float32(x)
    if np.isnan(x):
        return "-nan" if np.signbit(x) else "nan"
    if np.isinf(x):
        return "-inf" if x < 0 else "inf"
    sign = "-" if np.signbit(x) else ""
    ax = abs(x)
    if ax == 0:
        return sign + "0"
    sci = np.format_float_scientific(ax, unique=True, trim="-", exp_digits=2)  # d.ddde+XX, shortest round-trip digits
    mant, exp = sci.split("e")
    e = int(exp)
    digits = mant.replace(".", "")
    sci_str = mant + "e" + ("-" if e < 0 else "+") + f"{abs(e):02d}"
    if e >= 0:
        # an integer-valued float whose shortest digits end before the units: libstdc++ prints the EXACT integer there
        # (510307072, not 510307070), which has the same length as the zero-padded digits
        fixed = str(int(ax)) if len(digits) - 1 <= e else digits[: e + 1] + "." + digits[e + 1:]
    else:
        fixed = "0." + "0" * (-e - 1) + digits
    return sign + (fixed if len(fixed) <= len(sci_str) else sci_str)


def point_cloud_text(bgr: np.ndarray, xyz: np.ndarray, fmt: str) -> bytes:
    """The exact bytes writePCD / writePLY produce for the finite points of `xyz` (HxWx3 float32) coloured by `bgr`."""
    pts = xyz.reshape(-1, 3).astype(np.float32)
    col = bgr.reshape(-1, 3)
    keep = ~np.isposinf(pts).any(axis=1)  # == +infinity only, like the reference (:263-266)
    pts, col = pts[keep], col[keep]
    n = len(pts)
    if fmt == "pcd":
        out = ["# .PCD v0.7 - Point Cloud Data file format\n", "VERSION 0.7\n", "FIELDS x y z rgb\n", "SIZE 4 4 4 4\n",
               "TYPE F F F U\n", "COUNT 1 1 1 1\n", f"WIDTH {n}\n", "HEIGHT 1\n", "VIEWPOINT 0 0 0 1 0 0 0\n", f"POINTS {n}\n",
               "DATA ascii\n"]
        for p, c in zip(pts, col):
            rgb = int(c[2]) << 16 | int(c[1]) << 8 | int(c[0]) | 1 << 24
            out.append(f"{to_chars_f32(p[0])} {to_chars_f32(p[1])} {to_chars_f32(p[2])} {rgb}\n")
    else:
        out = ["ply\n", "format ascii 1.0\n", f"element vertex {n}\n", "property float x\n", "property float y\n",
               "property float z\n", "property uchar red\n", "property uchar green\n", "property uchar blue\n", "end_header\n"]
        for p, c in zip(pts, col):
            out.append(f"{to_chars_f32(p[0])} {to_chars_f32(p[1])} {to_chars_f32(p[2])} {int(c[2])} {int(c[1])} {int(c[0])}\n")
    return "".join(out).encode()
