"""Synthetic stereo inputs for the benchmark configs (synth_v1 and the
C4 rectify maps).  Pure numpy + scipy; deterministic for a given seed."""
from __future__ import annotations

import numpy as np
from scipy.ndimage import gaussian_filter


def _unit_noise(rng, shape, sigma):
    n = gaussian_filter(rng.standard_normal(shape).astype(np.float32), sigma, mode="wrap")
    return n / max(float(n.std()), 1e-6)


def synth_v1(H: int, W: int, Dmax: int, seed: int, gray: bool = False):
    """Layered scene with true occlusions.  Returns (left, right) uint8 HxWx3 (BGR order)."""
    rng = np.random.default_rng(seed)
    C = 1 if gray else 3

    def texture(h, w):
        off = rng.uniform(-40, 40, C).astype(np.float32)
        base = 128 + 5 * _unit_noise(rng, (h, w), 1.5) + 35 * _unit_noise(rng, (h, w), 10.0)
        return base[:, :, None] + off[None, None, :]

    left = np.zeros((H, W, C), np.float32)
    right = np.zeros((H, W, C), np.float32)
    # background: per-row disparity, texture wide enough for the shifted right view
    bg = texture(H, W + Dmax)
    drow = np.rint(0.10 * Dmax + 0.15 * Dmax * np.arange(H) / H).astype(int)
    xs = np.arange(W)
    for y in range(H):
        left[y] = bg[y, xs + 0]
        right[y] = bg[y, xs + drow[y]]
    K = 24
    rects = []
    for _ in range(K):
        w = int(rng.integers(W // 16, W // 4 + 1))
        h = int(rng.integers(H // 16, H // 4 + 1))
        x0 = int(rng.integers(0, W - w + 1))
        y0 = int(rng.integers(0, H - h + 1))
        d = int(rng.integers(int(np.ceil(0.2 * Dmax)), int(np.floor(0.9 * Dmax)) + 1))
        rects.append((d, x0, y0, w, h))
    rects.sort(key=lambda r: r[0])
    for d, x0, y0, w, h in rects:
        tex = texture(h, w)
        left[y0:y0 + h, x0:x0 + w] = tex
        xr0 = x0 - d
        lo, hi = max(xr0, 0), min(xr0 + w, W)
        if hi > lo:
            right[y0:y0 + h, lo:hi] = tex[:, lo - xr0:hi - xr0]
    out = []
    for view, img in enumerate((left, right)):
        nrng = np.random.default_rng([seed, view])
        img = img + nrng.normal(0, 0.7, img.shape).astype(np.float32)
        img = np.clip(np.rint(img), 0, 255).astype(np.uint8)
        if gray:
            img = np.repeat(img, 3, axis=2)
        out.append(np.ascontiguousarray(img))
    return out[0], out[1]


def synth_rectify_maps(H: int = 1024, W: int = 1280):
    """Float32 (mx, my) map pairs for the left and right halves of the C4 frame."""
    f, cx, cy = 1100.0, W / 2.0, H / 2.0
    ys, xs = np.mgrid[0:H, 0:W].astype(np.float64)
    u, v = (xs - cx) / f, (ys - cy) / f
    out = []
    for theta_deg, k1, tx, ty in ((0.3, -0.08, 3.25, -1.5), (-0.2, -0.07, -2.75, 2.25)):
        th = np.deg2rad(theta_deg)
        s = 1 + k1 * (u * u + v * v)
        mx = cx + f * s * (u * np.cos(th) - v * np.sin(th)) + tx
        my = cy + f * s * (u * np.sin(th) + v * np.cos(th)) + ty
        out.append((mx.astype(np.float32), my.astype(np.float32)))
    return out


def convert_maps_fixed(mx: np.ndarray, my: np.ndarray):
    """cv::convertMaps(mx, my, CV_16SC2): ix = rint(x*32) -> (ix>>5, iy>>5), (iy&31)*32 + (ix&31)."""
    ix = np.rint(mx.astype(np.float32) * np.float32(32)).astype(np.int64)
    iy = np.rint(my.astype(np.float32) * np.float32(32)).astype(np.int64)
    m1 = np.stack([np.clip(ix >> 5, -32768, 32767), np.clip(iy >> 5, -32768, 32767)], axis=2).astype(np.int16)
    m2 = ((iy & 31) * 32 + (ix & 31)).astype(np.uint16)
    return np.ascontiguousarray(m1), np.ascontiguousarray(m2)
