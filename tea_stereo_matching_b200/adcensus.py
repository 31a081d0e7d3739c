"""Python mirror of the reference's stereo::ADCensus operator, over the C-ABI.

Same names, argument meaning and error behaviour as the reference class
(include/stereo.h:388-422, source/ADCensus.cpp:298-407):

    m = ADCensus()                       # default: HSI, D = 0..64   (ADCensus.cpp:409-420)
    m.setMatchingStrategy(ColorModel.RGB, False, False)
    m.setMinMaxDisparity(0, 192)
    disparity = m.compute(left, right)   # HxWx3 uint8 BGR -> HxW float32

The reference throws ``std::string`` from the setters and the image check; here
those become ``ADCensusError`` with the same message text.  All compute goes
through libtsm_b200.so on a CUDA device -- there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import enum

import numpy as np

from . import _native as N


class ColorModel(enum.IntEnum):
    """stereo::ColorModel, include/stereo_utils.h:191-195"""

    RGB = 0
    HSI = 1


class ADCensusError(RuntimeError):
    """What the reference throws as std::string / std::runtime_error."""

    def __init__(self, message: str, status: int = N.TSM_E_ARG):
        super().__init__(message)
        self.status = status


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _as_bgr(img, what="image"):
    a = np.asarray(img)
    if a.size == 0:
        raise ADCensusError("[ADCensus] Image error.")
    if a.dtype != np.uint8 or a.ndim != 3 or a.shape[2] != 3:
        raise ADCensusError(f"[ADCensus] Image error ({what} must be HxWx3 uint8, CV_8UC3).")
    if a.strides[2] != 1 or a.strides[1] != 3:
        a = np.ascontiguousarray(a)
    return a


class Context:
    """One tsm_ctx: one CUDA device + one stream.  Not thread-safe (like the reference's Impl)."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self._lib = N.lib()
        h = C.c_void_p()
        rc = self._lib.tsm_create_on_stream(device, C.c_void_p(stream) if stream else None, C.byref(h))
        if rc != N.TSM_OK:
            raise ADCensusError(self._lib.tsm_last_error(None).decode(), rc)
        self.handle = h
        self.device = device

    def check(self, rc: int) -> None:
        if rc != N.TSM_OK:
            raise ADCensusError(self._lib.tsm_last_error(self.handle).decode(), rc)

    def close(self) -> None:
        if getattr(self, "handle", None):
            self._lib.tsm_destroy(self.handle)
            self.handle = None

    def synchronize(self) -> None:
        self.check(self._lib.tsm_synchronize(self.handle))

    @property
    def launch_count(self) -> int:
        return int(self._lib.tsm_launch_count(self.handle))

    def set_profiling(self, on: bool) -> None:
        self.check(self._lib.tsm_set_profiling(self.handle, int(on)))

    def stage_times(self) -> dict:
        """Device time per stage of the last profiled compute, ms.  Keys with a '/' ("aggregate/h_norm") are single
        launches inside the stage named before the slash (summed over repeats; `<key>#n` counts them)."""
        n = C.c_int(96)
        names = (C.c_char_p * 96)()
        ms = (C.c_float * 96)()
        self.check(self._lib.tsm_get_stage_times(self.handle, C.byref(n), names, ms))
        out: dict = {}
        for i in range(n.value):
            k = names[i].decode()
            out[k] = out.get(k, 0.0) + float(ms[i])
            if "/" in k:
                out[k + "#n"] = out.get(k + "#n", 0) + 1
        return out

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ADCensus:
    """Drop-in mirror of stereo::ADCensus."""

    def __init__(self, device: int = 0, stream: int | None = None, context: Context | None = None):
        # ADCensusImpl::ADCensusImpl, ADCensus.cpp:409-420
        self._min = 0
        self._max = 64
        self._model = ColorModel.HSI
        self._roi = False
        self._mask = False
        self._offset = 0
        self._ctx = context
        self._device = device
        self._stream = stream
        self.last_shape = (0, 0)  # geometry of the map of the last compute / enqueue (consumers.py)

    # -- reference API ----------------------------------------------------
    def setMinMaxDisparity(self, minDisparity: int, maxDisparity: int) -> None:
        if minDisparity * maxDisparity < 0 or minDisparity >= maxDisparity:  # ADCensus.cpp:309
            raise ADCensusError("[ADCensus] Set MinMaxDisparity error.")
        self._min, self._max = int(minDisparity), int(maxDisparity)

    def setMatchingStrategy(self, colorModel: ColorModel = ColorModel.RGB, roiMatching: bool = False,
                            maskMatching: bool = False) -> None:
        self._model = ColorModel(colorModel)
        self._roi, self._mask = bool(roiMatching), bool(maskMatching)

    def setOffset(self, offset: int) -> None:
        if offset < 0:  # ADCensus.cpp:325
            raise ADCensusError("[ADCensus] Offset must be positive.")
        self._offset = int(offset)

    def compute(self, leftImage, rightImage) -> np.ndarray:
        """ADCensus::compute: returns the CV_32FC1 disparity of the left view."""
        left, right = self._check_pair(leftImage, rightImage)
        H, W, _ = left.shape
        out = np.empty((H, W), np.float32)
        ctx = self.context
        ctx.check(ctx._lib.tsm_adcensus_compute(ctx.handle, C.byref(self._config()), _ptr(left), left.strides[0],
                                                 _ptr(right), right.strides[0], H, W, _ptr(out), out.strides[0]))
        self.last_shape = (H, W)
        return out

    # -- extras over the reference ------------------------------------------
    def enqueue(self, leftImage, rightImage) -> None:
        left, right = self._check_pair(leftImage, rightImage)
        H, W, _ = left.shape
        self._pending_shape = self.last_shape = (H, W)
        ctx = self.context
        ctx.check(ctx._lib.tsm_adcensus_enqueue(ctx.handle, C.byref(self._config()), _ptr(left), left.strides[0],
                                                 _ptr(right), right.strides[0], H, W))

    def wait(self, out: np.ndarray | None = None) -> np.ndarray:
        H, W = self._pending_shape
        if out is None:
            out = np.empty((H, W), np.float32)
        elif (not isinstance(out, np.ndarray) or out.dtype != np.float32 or out.shape != (H, W) or out.strides[1] != 4
              or out.strides[0] < 4 * W or not out.flags.writeable):
            # the C side copies H rows of W floats at out.strides[0]: anything else would overrun the buffer
            raise ADCensusError(f"[ADCensus] disparity buffer error (need a writeable float32 array of shape ({H}, {W}) with contiguous rows).")
        ctx = self.context
        ctx.check(ctx._lib.tsm_adcensus_wait(ctx.handle, _ptr(out), out.strides[0]))
        return out

    IN_FLIGHT = 4  # pairs in flight per device of computeBatch (measured: 2 -> 17.1, 3 -> 16.7, 4 -> 16.5, 6 -> 16.5 ms per 1080p pair)

    def computeBatch(self, leftImages, rightImages, devices=None) -> list:
        """ADCensus::compute(vector, vector, vector&) of the C++ facade (cpp/stereo.h; the reference's batched signature
        style, include/stereo.h:381): pair i runs on devices[i % N], one worker thread per device, IN_FLIGHT contexts each.
        devices=None: this object's device; devices=-1: every visible device.  Sharding never changes a pair's bits."""
        import threading

        if len(leftImages) != len(rightImages):
            raise ADCensusError("[ADCensus] Image error.")
        pairs = [self._check_pair(l, r) for l, r in zip(leftImages, rightImages)]
        n = len(pairs)
        if n == 0:
            return []
        if devices is None:
            devices = [self._device]
        elif devices == -1:
            cnt = C.c_int32(0)
            if N.lib().tsm_device_count(C.byref(cnt)) != N.TSM_OK or cnt.value < 1:
                raise ADCensusError(N.lib().tsm_last_error(None).decode(), N.TSM_E_CUDA)
            devices = list(range(min(cnt.value, n)))
        devices = list(devices)
        pool = self.__dict__.setdefault("_batch_pool", {})
        out: list = [None] * n
        errors: list = []

        def worker(w: int) -> None:
            ms = pool.get(devices[w])
            if ms is None:
                ms = pool[devices[w]] = [ADCensus(device=devices[w]) for _ in range(self.IN_FLIGHT)]
            for m in ms:
                m._min, m._max, m._model, m._roi, m._mask, m._offset = (self._min, self._max, self._model, self._roi,
                                                                        self._mask, self._offset)
            mine = list(range(w, n, len(devices)))
            K = len(ms)
            slot: list = [None] * K
            try:
                for j in range(len(mine) + K):
                    k = j % K
                    if slot[k] is not None:
                        i, slot[k] = slot[k], None
                        out[i] = ms[k].wait()
                    if j < len(mine):
                        ms[k].enqueue(*pairs[mine[j]])
                        slot[k] = mine[j]
            except Exception as e:  # leave no context with an un-waited pair, then report
                for k in range(K):
                    if slot[k] is not None:
                        try:
                            ms[k].wait()
                        except Exception:
                            pass
                errors.append(e)

        if len(devices) == 1:
            worker(0)
        else:
            ths = [threading.Thread(target=worker, args=(w,)) for w in range(len(devices))]
            [t.start() for t in ths]
            [t.join() for t in ths]
        if errors:
            raise errors[0]
        return out

    def compute_device(self, d_left: int, d_right: int, H: int, W: int, d_out: int) -> None:
        """Device-resident form: raw device pointers (packed BGR in, packed float out), async on the ctx stream."""
        ctx = self.context
        ctx.check(ctx._lib.tsm_adcensus_compute_device(ctx.handle, C.byref(self._config()), C.c_void_p(d_left),
                                                        C.c_void_p(d_right), H, W, C.c_void_p(d_out)))
        self.last_shape = (H, W)

    @property
    def context(self) -> Context:
        if self._ctx is None:
            self._ctx = Context(self._device, self._stream)
        return self._ctx

    def _config(self) -> N.Config:
        return N.Config(self._min, self._max, int(self._model), int(self._roi), int(self._mask), self._offset)

    @staticmethod
    def _check_pair(leftImage, rightImage):
        if leftImage is None or rightImage is None:
            raise ADCensusError("[ADCensus] Image error.")
        left, right = _as_bgr(leftImage, "left"), _as_bgr(rightImage, "right")
        if left.shape != right.shape:  # ADCensus.cpp:332
            raise ADCensusError("[ADCensus] Image error.")
        return left, right


class StageRunner:
    """Parity harness over tsm_stage_begin / tsm_stage_run / tsm_tap / tsm_poke (tests only)."""

    def __init__(self, left, right, max_disparity: int, device: int = 0, model: "ColorModel | None" = None,
                 min_disparity: int = 0):
        self.ctx = Context(device)
        self.left, self.right = _as_bgr(left), _as_bgr(right)
        self.H, self.W, _ = self.left.shape
        self.Dn = max_disparity - min_disparity + 1  # cost planes (ADCensus.cpp:345)
        self.cfg = N.Config(min_disparity, max_disparity, int(ColorModel.RGB if model is None else model), 0, 0, 0)
        L = self.ctx._lib
        self.ctx.check(L.tsm_stage_begin(self.ctx.handle, C.byref(self.cfg), _ptr(self.left), self.left.strides[0],
                                         _ptr(self.right), self.right.strides[0], self.H, self.W))
        self.Dp = int(L.tsm_volume_pitch(self.ctx.handle))

    def run(self, mask: int, arg: int = -1) -> None:
        self.ctx.check(self.ctx._lib.tsm_stage_run(self.ctx.handle, mask, arg))

    def _tap_raw(self, buf: int, dtype, shape):
        a = np.empty(shape, dtype)
        self.ctx.check(self.ctx._lib.tsm_tap(self.ctx.handle, buf, _ptr(a), a.nbytes))
        return a

    def volume(self, view: int) -> np.ndarray:
        v = self._tap_raw(N.BUF_VOL_LEFT + view, np.float32, (self.H, self.W, self.Dp))
        return np.ascontiguousarray(v[:, :, : self.Dn])

    def set_volume(self, view: int, vol: np.ndarray) -> None:
        v = np.zeros((self.H, self.W, self.Dp), np.float32)
        v[:, :, : self.Dn] = vol
        self.ctx.check(self.ctx._lib.tsm_poke(self.ctx.handle, N.BUF_VOL_LEFT + view, _ptr(v), v.nbytes))

    def image(self, view: int) -> np.ndarray:
        """The 3-channel image the matching stages see (BGR, or H,S,I after the HSI preprocessing)."""
        v = self._tap_raw(N.BUF_IMG4_LEFT + view, np.uint8, (self.H, self.W, 4))
        return np.ascontiguousarray(v[:, :, :3])

    def arms(self, view: int) -> np.ndarray:
        return self._tap_raw(N.BUF_ARMS_LEFT + view, np.uint8, (self.H, self.W, 4))

    def census(self, view: int) -> np.ndarray:
        return self._tap_raw(N.BUF_CENSUS_LEFT + view, np.uint64, (6, self.H, self.W))

    def wta(self, view: int) -> np.ndarray:
        return self._tap_raw(N.BUF_WTA_LEFT + view, np.int32, (self.H, self.W))

    def set_wta(self, view: int, d: np.ndarray) -> None:
        d = np.ascontiguousarray(d, np.int32)
        self.ctx.check(self.ctx._lib.tsm_poke(self.ctx.handle, N.BUF_WTA_LEFT + view, _ptr(d), d.nbytes))

    def disp(self) -> np.ndarray:
        return self._tap_raw(N.BUF_DISP, np.int32, (self.H, self.W))

    def set_disp(self, d: np.ndarray) -> None:
        d = np.ascontiguousarray(d, np.int32)
        self.ctx.check(self.ctx._lib.tsm_poke(self.ctx.handle, N.BUF_DISP, _ptr(d), d.nbytes))

    def edges(self) -> np.ndarray:
        return self._tap_raw(N.BUF_EDGES, np.uint8, (self.H, self.W))

    def final(self) -> np.ndarray:
        return self._tap_raw(N.BUF_FINAL, np.float32, (self.H, self.W))

    def close(self) -> None:
        self.ctx.close()
