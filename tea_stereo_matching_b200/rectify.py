"""Python mirror of stereo::EpipolarRectify (include/stereo.h:254-296,
source/EpipolarRectify.cpp) and of the EpipolarRectifyMap value type
(include/stereo_utils.h:109-148), over the C-ABI (tsm_remap / tsm_rectify_stereo).

Error behaviour follows the reference: loadEpipolarRectifyMap raises on empty maps
(EpipolarRectify.cpp:35-40); the rectify overloads log and return None, leaving the
outputs untouched, when the maps or the image are empty (:48-57, :70-79, :89-98).
"""
from __future__ import annotations

import ctypes as C
import sys
from dataclasses import dataclass

import numpy as np

from . import _native as N
from .adcensus import ADCensusError, Context, _as_bgr, _ptr


@dataclass
class CameraIntrinsic:
    """stereo::CameraIntrinsic (include/stereo_utils.h): 3x3 K and the OpenCV distortion vector."""

    intrinsic_matrix: np.ndarray | None = None
    distortion_coefficients: np.ndarray | None = None

    def empty(self) -> bool:
        return self.intrinsic_matrix is None or np.asarray(self.intrinsic_matrix).size == 0


@dataclass
class StereoPair:
    left: CameraIntrinsic | None = None
    right: CameraIntrinsic | None = None


def initUndistortRectifyMap(cameraMatrix, distCoeffs, R, newCameraMatrix, size: tuple[int, int], context: Context | None = None):
    """cv::initUndistortRectifyMap(cameraMatrix, distCoeffs, R, newCameraMatrix, size, CV_16SC2, map1, map2) on the
    device (csrc/k_rectify.cu); size = (width, height).  Returns (map1 CV_16SC2, map2 CV_16UC1)."""
    ctx = context or Context(0)
    W, H = int(size[0]), int(size[1])
    dbl = C.POINTER(C.c_double)
    K = np.ascontiguousarray(cameraMatrix, np.float64).reshape(3, 3)
    D = np.zeros(0) if distCoeffs is None else np.ascontiguousarray(distCoeffs, np.float64).reshape(-1)
    Rm = None if R is None or np.asarray(R).size == 0 else np.ascontiguousarray(R, np.float64).reshape(3, 3)
    P = None if newCameraMatrix is None or np.asarray(newCameraMatrix).size == 0 else np.ascontiguousarray(newCameraMatrix, np.float64)
    if P is not None and P.shape not in ((3, 3), (3, 4)):
        raise ADCensusError("[EpipolarRectify] newCameraMatrix must be 3x3 or 3x4")
    map1, map2 = np.empty((H, W, 2), np.int16), np.empty((H, W), np.uint16)
    ctx.check(ctx._lib.tsm_init_undistort_rectify_map(
        ctx.handle, K.ctypes.data_as(dbl), D.ctypes.data_as(dbl) if D.size else None, int(D.size),
        Rm.ctypes.data_as(dbl) if Rm is not None else None, P.ctypes.data_as(dbl) if P is not None else None,
        0 if P is None else P.shape[1], H, W, _ptr(map1), map1.strides[0], _ptr(map2), map2.strides[0]))
    return map1, map2


@dataclass
class EpipolarRectifyMap:
    """stereo::EpipolarRectifyMap (include/stereo_utils.h:109-148, source/stereo_utils.cpp:134-174)."""

    R1: np.ndarray | None = None
    R2: np.ndarray | None = None
    P1: np.ndarray | None = None
    P2: np.ndarray | None = None
    map00: np.ndarray | None = None
    map01: np.ndarray | None = None
    map10: np.ndarray | None = None
    map11: np.ndarray | None = None

    def empty(self) -> bool:  # stereo_utils.cpp:171-174
        return any(m is None or np.asarray(m).size == 0 for m in (self.map00, self.map01, self.map10, self.map11))

    def compute(self, intrinsic: StereoPair, imgsz: tuple[int, int], context: Context | None = None) -> None:
        """EpipolarRectifyMap::compute (stereo_utils.cpp:157-169): the four CV_16SC2 / CV_16UC1 maps from
        K, D of both cameras and this object's R1, R2, P1, P2; a no-op when an intrinsic is empty."""
        if intrinsic.left is None or intrinsic.right is None or intrinsic.left.empty() or intrinsic.right.empty():
            return
        ctx = context or Context(0)
        self.map00, self.map01 = initUndistortRectifyMap(intrinsic.left.intrinsic_matrix, intrinsic.left.distortion_coefficients,
                                                         self.R1, self.P1, imgsz, ctx)
        self.map10, self.map11 = initUndistortRectifyMap(intrinsic.right.intrinsic_matrix, intrinsic.right.distortion_coefficients,
                                                         self.R2, self.P2, imgsz, ctx)


def _log_error(msg: str) -> None:
    print(f"[ERROR] {msg}", file=sys.stderr)


def _map_kind(m1: np.ndarray, m2: np.ndarray) -> int:
    if m1.dtype == np.int16 and m1.ndim == 3 and m1.shape[2] == 2 and m2.dtype == np.uint16:
        return N.MAP_FIXED
    if m1.dtype == np.float32 and m2.dtype == np.float32 and m1.ndim == 2:
        return N.MAP_FLOAT
    raise ADCensusError("[EpipolarRectify] unsupported map types (need CV_16SC2+CV_16UC1 or CV_32FC1 x2)")


# Process-wide id of the map CONTENT handed to the C-ABI (tsm.h: map_generation): every loadEpipolarRectifyMap takes a
# new one, so a device-side copy is never reused for different maps that happen to live at the same host addresses.
import itertools

_map_generation = itertools.count(1)


class EpipolarRectify:
    def __init__(self, rectifyMap: EpipolarRectifyMap | None = None, imgsz: tuple[int, int] | None = None,
                 device: int = 0, context: Context | None = None):
        self._map = EpipolarRectifyMap()
        self._imgsz = (0, 0)  # (width, height), cv::Size order
        self._ctx = context
        self._device = device
        if rectifyMap is not None:
            self.loadEpipolarRectifyMap(rectifyMap, imgsz)

    @property
    def context(self) -> Context:
        if self._ctx is None:
            self._ctx = Context(self._device)
        return self._ctx

    def loadEpipolarRectifyMap(self, rectifyMap: EpipolarRectifyMap, imgsz: tuple[int, int]) -> None:
        if rectifyMap is None or rectifyMap.empty():
            raise RuntimeError("stereo params is empty, please load it first")  # EpipolarRectify.cpp:37-39
        m = EpipolarRectifyMap(**{k: (None if v is None else np.ascontiguousarray(v)) for k, v in vars(rectifyMap).items()})
        self._kind = _map_kind(m.map00, m.map01)
        if _map_kind(m.map10, m.map11) != self._kind:
            raise ADCensusError("[EpipolarRectify] left and right maps must have the same type")
        self._map = m
        self._gen = next(_map_generation)
        self._imgsz = (int(imgsz[0]), int(imgsz[1]))

    # rectify(left, right) -> (rectLeft, rectRight)         EpipolarRectify.cpp:87-101
    # rectify(stereo)      -> (rectLeft, rectRight)         EpipolarRectify.cpp:68-85
    def rectify(self, *images):
        if self._map.empty():
            _log_error("Stereo epipolar rectify params is empty, please load it first.")
            return None
        if len(images) == 2:
            left, right = images
            if left is None or right is None or np.asarray(left).size == 0 or np.asarray(right).size == 0:
                _log_error("Left or Right image is empty.")
                return None
            return self._remap(_as_bgr(left), 0), self._remap(_as_bgr(right), 1)
        if len(images) != 1:
            raise TypeError("rectify(stereoImage) or rectify(leftImage, rightImage)")
        stereo = images[0]
        if stereo is None or np.asarray(stereo).size == 0:
            _log_error("Stereo image is empty.")
            return None
        stereo = _as_bgr(stereo, "stereo")
        W, H = self._imgsz
        if stereo.shape[0] < H or stereo.shape[1] < 2 * W:
            raise ADCensusError("[EpipolarRectify] stereo image smaller than 2*imgsz")
        mh, mw = self._map.map00.shape[:2]
        if (mh, mw) != (H, W):
            # general (map size != crop size) case: two independent remaps of the cropped halves
            return self._remap(stereo[:H, :W], 0), self._remap(stereo[:H, W:2 * W], 1)
        left = np.empty((H, W, 3), np.uint8)
        right = np.empty((H, W, 3), np.uint8)
        ctx, m = self.context, self._map
        ctx.check(ctx._lib.tsm_rectify_stereo(ctx.handle, _ptr(stereo), stereo.strides[0], H, W, _ptr(m.map00), _ptr(m.map01),
                                               _ptr(m.map10), _ptr(m.map11), self._kind, self._gen, _ptr(left), left.strides[0],
                                               _ptr(right), right.strides[0]))
        return left, right

    def rectifyStereo(self, stereoImage):
        """rectify(stereoImage, rectifiedStereoImage): hconcat of the two rectified halves (EpipolarRectify.cpp:46-66)."""
        r = self.rectify(stereoImage)
        return None if r is None else np.concatenate(r, axis=1)

    def rectify_adcensus(self, stereoImage, matcher) -> np.ndarray:
        """Fused rectify -> ADCensus (BASELINE config C4); `matcher` is an ADCensus holding the disparity setup."""
        if self._map.empty():
            raise RuntimeError("stereo params is empty, please load it first")
        stereo = _as_bgr(stereoImage, "stereo")
        W, H = self._imgsz
        out = np.empty((H, W), np.float32)
        ctx, m = matcher.context, self._map
        ctx.check(ctx._lib.tsm_rectify_adcensus(ctx.handle, C.byref(matcher._config()), _ptr(stereo), stereo.strides[0], H, W,
                                                 _ptr(m.map00), _ptr(m.map01), _ptr(m.map10), _ptr(m.map11), self._kind,
                                                 self._gen, _ptr(out), out.strides[0]))
        return out

    def rectify_adcensus_enqueue(self, stereoImage, matcher) -> None:
        """Asynchronous rectify_adcensus: collect the map with matcher.wait()."""
        if self._map.empty():
            raise RuntimeError("stereo params is empty, please load it first")
        stereo = _as_bgr(stereoImage, "stereo")
        W, H = self._imgsz
        ctx, m = matcher.context, self._map
        ctx.check(ctx._lib.tsm_rectify_adcensus_enqueue(ctx.handle, C.byref(matcher._config()), _ptr(stereo), stereo.strides[0], H, W,
                                                         _ptr(m.map00), _ptr(m.map01), _ptr(m.map10), _ptr(m.map11), self._kind,
                                                         self._gen))
        matcher._pending_shape = matcher.last_shape = (H, W)

    def rectify_adcensus_device(self, d_stereo: int, sstep: int, matcher, d_out: int) -> None:
        """Device-resident form of rectify_adcensus: raw device pointers, async on the matcher's stream."""
        W, H = self._imgsz
        ctx, m = matcher.context, self._map
        ctx.check(ctx._lib.tsm_rectify_adcensus_device(ctx.handle, C.byref(matcher._config()), C.c_void_p(d_stereo), sstep, H, W,
                                                        _ptr(m.map00), _ptr(m.map01), _ptr(m.map10), _ptr(m.map11), self._kind,
                                                        self._gen, C.c_void_p(d_out)))
        matcher.last_shape = (H, W)

    def _remap(self, src: np.ndarray, which: int) -> np.ndarray:
        m1, m2 = (self._map.map00, self._map.map01) if which == 0 else (self._map.map10, self._map.map11)
        H, W = m1.shape[:2]
        dst = np.empty((H, W, 3), np.uint8)
        ctx = self.context
        ctx.check(ctx._lib.tsm_remap(ctx.handle, _ptr(src), src.strides[0], src.shape[0], src.shape[1], _ptr(m1), _ptr(m2),
                                      self._kind, self._gen, H, W, _ptr(dst), dst.strides[0]))
        return dst


# ---- stereo::StereoParams (include/stereo_utils.h, source/stereo_utils.cpp:176-232) --------------------
def _read_opencv_yaml(path: str) -> dict:
    """Reads the subset of OpenCV FileStorage YAML the reference writes: `key: !!opencv-matrix` blocks
    (rows, cols, dt, data) and flow sequences `key: [ a, b ]`."""
    import re

    with open(path, "r") as f:
        text = f.read()
    out: dict = {}
    for m in re.finditer(r"^(\w+):\s*!!opencv-matrix\s*\n\s*rows:\s*(\d+)\s*\n\s*cols:\s*(\d+)\s*\n\s*dt:\s*\"?(\w+)\"?\s*\n\s*data:\s*\[(.*?)\]",
                         text, re.S | re.M):
        key, rows, cols, dt, data = m.group(1), int(m.group(2)), int(m.group(3)), m.group(4), m.group(5)
        vals = [float(v) for v in data.replace("\n", " ").split(",") if v.strip()]
        np_dt = {"d": np.float64, "f": np.float32, "i": np.int32, "s": np.int16, "w": np.uint16, "u": np.uint8}[dt[-1]]
        ch = int(dt[:-1]) if len(dt) > 1 else 1
        a = np.array(vals, np_dt)
        out[key] = a.reshape(rows, cols, ch) if ch > 1 else a.reshape(rows, cols)
    for m in re.finditer(r"^(\w+):\s*\[([^\]]*)\]\s*$", text, re.M):
        out.setdefault(m.group(1), [float(v) for v in m.group(2).split(",") if v.strip()])
    return out


class StereoParams:
    """Loads a stereo calibration YAML like the reference and computes the rectify maps on the device."""

    def __init__(self, ymlFilePath: str | None = None, context: Context | None = None):
        self.intrinsic = StereoPair(CameraIntrinsic(), CameraIntrinsic())
        self.E = self.F = self.R = self.T = self.Q = None
        self.map = EpipolarRectifyMap()
        self.rectified_f = self.rectified_cx = self.rectified_cy = self.baseline = 0.0
        self.imgsz = (0, 0)
        self._ctx = context
        if ymlFilePath is not None:
            self.loadYAMLFile(ymlFilePath)

    def loadYAMLFile(self, ymlFilePath: str) -> None:
        if not ymlFilePath:
            raise ValueError("Stereo YAML file path is empty.")  # std::invalid_argument, stereo_utils.cpp:187-192
        try:
            y = _read_opencv_yaml(ymlFilePath)
        except OSError:
            raise RuntimeError("Cannot open stereo yml file.") from None  # :197-203
        self.intrinsic = StereoPair(CameraIntrinsic(y.get("leftK"), y.get("leftD")), CameraIntrinsic(y.get("rightK"), y.get("rightD")))
        self.E, self.F, self.R, self.T, self.Q = (y.get(k) for k in ("E", "F", "R", "T", "Q"))
        self.map.R1, self.map.R2, self.map.P1, self.map.P2 = (y.get(k) for k in ("R1", "R2", "P1", "P2"))
        if "imgsz" in y:
            sz = np.asarray(y["imgsz"]).reshape(-1)
            self.imgsz = (int(sz[0]), int(sz[1]))
        if self.Q is None:
            return
        Q = np.asarray(self.Q, np.float64)
        self.rectified_f = float(np.float32(Q[2, 3]))      # :222-225
        self.rectified_cx = float(np.float32(-Q[0, 3]))
        self.rectified_cy = float(np.float32(-Q[1, 3]))
        self.baseline = float(np.float32(1.0) / np.float32(Q[3, 2]))
        self.map.compute(self.intrinsic, self.imgsz, self._ctx)

    def empty(self) -> bool:  # :234-238
        return (self.intrinsic.left.empty() or self.intrinsic.right.empty() or self.R is None or self.T is None
                or self.map.empty() or self.Q is None)
