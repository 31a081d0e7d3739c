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
class EpipolarRectifyMap:
    """R1,R2,P1,P2 are carried for API parity; only the four maps are used on this path."""

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


def _log_error(msg: str) -> None:
    print(f"[ERROR] {msg}", file=sys.stderr)


def _map_kind(m1: np.ndarray, m2: np.ndarray) -> int:
    if m1.dtype == np.int16 and m1.ndim == 3 and m1.shape[2] == 2 and m2.dtype == np.uint16:
        return N.MAP_FIXED
    if m1.dtype == np.float32 and m2.dtype == np.float32 and m1.ndim == 2:
        return N.MAP_FLOAT
    raise ADCensusError("[EpipolarRectify] unsupported map types (need CV_16SC2+CV_16UC1 or CV_32FC1 x2)")


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
        self._imgsz = (int(imgsz[0]), int(imgsz[1]))
        if self._ctx is not None:
            self._ctx._lib.tsm_invalidate_maps(self._ctx.handle)

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
                                               _ptr(m.map10), _ptr(m.map11), self._kind, _ptr(left), left.strides[0],
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
                                                 _ptr(out), out.strides[0]))
        return out

    def _remap(self, src: np.ndarray, which: int) -> np.ndarray:
        m1, m2 = (self._map.map00, self._map.map01) if which == 0 else (self._map.map10, self._map.map11)
        H, W = m1.shape[:2]
        dst = np.empty((H, W, 3), np.uint8)
        ctx = self.context
        ctx.check(ctx._lib.tsm_remap(ctx.handle, _ptr(src), src.strides[0], src.shape[0], src.shape[1], _ptr(m1), _ptr(m2),
                                      self._kind, H, W, _ptr(dst), dst.strides[0]))
        return dst
