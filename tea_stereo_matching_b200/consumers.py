"""Consumers of the disparity map -- host-side mirror of the reference's free functions
(source/stereo.cpp:75-202) on top of the C-ABI (include/tsm.h, csrc/k_consumers.cu).

Same names and argument meaning as the reference; the cv::Mat outputs become return values.  Every function
takes either a host disparity map (HxW float32) or an `ADCensus` matcher: then the map of the matcher's last
`compute` / `wait` is consumed where it lies in device memory, without a D2H / H2D round trip.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as N
from .adcensus import ADCensus, ADCensusError, Context, _ptr


def _source(disparity, context: Context | None):
    """-> (context, pointer or None, row stride, H, W, keep-alive)"""
    if isinstance(disparity, ADCensus):
        H, W = disparity.last_shape
        return disparity.context, None, 0, H, W, None
    d = np.ascontiguousarray(disparity, np.float32)
    if d.ndim != 2:
        raise ValueError("disparity must be a HxW float32 map")
    ctx = context or _default_context()
    return ctx, d.ctypes.data_as(C.c_void_p), d.strides[0], d.shape[0], d.shape[1], d


_ctx: Context | None = None


def _default_context() -> Context:
    global _ctx
    if _ctx is None:
        _ctx = Context(0)
    return _ctx


def JETColorMap() -> np.ndarray:
    """stereo::JETColorMap() (stereo.cpp:75-93): 1x256 BGR table."""
    t = np.empty((1, 256, 3), np.uint8)
    N.lib().tsm_jet_colormap(t.ctypes.data_as(C.c_void_p))
    return t


def applyColorMap(src, minVal: float | None = None, maxVal: float | None = None, colorMap: np.ndarray | None = None,
                  context: Context | None = None) -> np.ndarray:
    """stereo::applyColorMap(src, dst, colorMap) / (src, dst, minVal, maxVal, colorMap) (stereo.cpp:95-137)."""
    ctx, ptr, step, H, W, keep = _source(src, context)
    auto = minVal is None or maxVal is None
    cm = None
    if colorMap is not None:
        cm = np.ascontiguousarray(colorMap, np.uint8).reshape(-1)
        if cm.size != 768:
            raise ValueError("colorMap must hold 256 BGR entries")
    dst = np.empty((H, W, 3), np.uint8)
    ctx.check(ctx._lib.tsm_apply_colormap(ctx.handle, ptr, step, H, W, 1 if auto else 0, 0.0 if auto else float(minVal),
                                          0.0 if auto else float(maxVal), None if cm is None else cm.ctypes.data_as(C.c_void_p),
                                          dst.ctypes.data_as(C.c_void_p), dst.strides[0]))
    return dst


def reprojectToDepth(disparity, focalLength: float, baseline: float, context: Context | None = None) -> np.ndarray:
    """stereo::reprojectToDepth(disparity, focalLength, baseline, depth) (stereo.cpp:139-151)."""
    ctx, ptr, step, H, W, keep = _source(disparity, context)
    depth = np.empty((H, W), np.float32)
    ctx.check(ctx._lib.tsm_reproject_to_depth(ctx.handle, ptr, step, H, W, float(focalLength), float(baseline),
                                              depth.ctypes.data_as(C.c_void_p), depth.strides[0]))
    return depth


def reprojectTo3D(disparity, *args, context: Context | None = None) -> np.ndarray:
    """stereo::reprojectTo3D(disparity, focalLength, baseline, cx, cy, XYZ) (stereo.cpp:153-172) or
    stereo::reprojectTo3D(disparity, Q, XYZ) with a 4x4 Q (stereo.cpp:174-202).  Returns HxWx3 float32."""
    ctx, ptr, step, H, W, keep = _source(disparity, context)
    xyz = np.empty((H, W, 3), np.float32)
    out = xyz.ctypes.data_as(C.c_void_p)
    if len(args) == 1:
        Q = np.ascontiguousarray(args[0], np.float64)
        if Q.shape != (4, 4):
            raise ValueError("Q must be 4x4")
        ctx.check(ctx._lib.tsm_reproject_to_3d_q(ctx.handle, ptr, step, H, W, Q.ctypes.data_as(C.POINTER(C.c_double)), out,
                                                 xyz.strides[0]))
    elif len(args) == 4:
        f, b, cx, cy = (float(a) for a in args)
        ctx.check(ctx._lib.tsm_reproject_to_3d(ctx.handle, ptr, step, H, W, f, b, cx, cy, out, xyz.strides[0]))
    else:
        raise TypeError("reprojectTo3D(disparity, focalLength, baseline, cx, cy) or reprojectTo3D(disparity, Q)")
    return xyz


def _write_cloud(RGBImage, XYZPoints, path: str, fmt: int) -> int:
    if RGBImage is None or XYZPoints is None or not path or np.asarray(RGBImage).size == 0 or np.asarray(XYZPoints).size == 0:
        print("[ERROR] Empty input.")  # the reference logs and returns (stereo.cpp:252-256)
        return 0
    rgb = np.ascontiguousarray(RGBImage, np.uint8)
    xyz = np.ascontiguousarray(XYZPoints, np.float32)
    if rgb.ndim != 3 or rgb.shape[2] != 3 or xyz.shape != rgb.shape:
        raise ADCensusError("writePointCloud: RGBImage must be HxWx3 uint8 and XYZPoints HxWx3 float32 of the same size")
    L = N.lib()
    n = C.c_size_t(0)
    rc = L.tsm_write_point_cloud(_ptr(rgb), rgb.strides[0], _ptr(xyz), xyz.strides[0], rgb.shape[0], rgb.shape[1],
                                 str(path).encode(), fmt, C.byref(n))
    if rc != N.TSM_OK:
        raise ADCensusError(L.tsm_last_error(None).decode(), rc)
    return int(n.value)


def writePointCloudToPCD(RGBImage, XYZPoints, pcdPath: str) -> int:
    """stereo::writePointCloudToPCD (stereo.cpp:250-278): ASCII PCD v0.7 of the finite points; returns the point count."""
    return _write_cloud(RGBImage, XYZPoints, pcdPath, 0)


def writePointCloudToPLY(RGBImage, XYZPoints, plyPath: str) -> int:
    """stereo::writePointCloudToPLY (stereo.cpp:328-356): ASCII PLY of the finite points; returns the point count."""
    return _write_cloud(RGBImage, XYZPoints, plyPath, 1)
