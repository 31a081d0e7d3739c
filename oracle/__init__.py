"""oracle/ -- TEST INFRASTRUCTURE, not product code.

Python loaders for the two CPU checkers of the ADCensus path:

* ``port``  -- oracle/liboracle.so, our plain-C restatement (adcensus_oracle.c, cvport.c).
* ``ref``   -- oracle/_ref/libadcensus_ref.so, the UNMODIFIED reference
  /root/reference/source/ADCensus.cpp compiled against oracle/ref_shim (ref_driver.cpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  The product (tea_stereo_matching_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
PORT_SO = HERE / "liboracle.so"
REF_SO = HERE / "_ref" / "libadcensus_ref.so"
REFERENCE_ROOT = Path(os.environ.get("TSM_REFERENCE_ROOT", "/root/reference"))


def build(ref: bool = True, port: bool = True) -> None:
    """Compile the checkers (make).  `ref` is skipped when /root/reference is absent."""
    if port:
        subprocess.run(["make", "-s", "-C", str(HERE), "port"], check=True)
    if ref and (REFERENCE_ROOT / "source" / "ADCensus.cpp").exists():
        subprocess.run(["make", "-s", "-C", str(HERE), "ref", f"REF={REFERENCE_ROOT}"], check=True)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class _Taps(C.Structure):
    _fields_ = [
        ("vol_init", C.c_void_p * 2),
        ("vol_agg", C.c_void_p * 2),
        ("vol_scan", C.c_void_p * 2),
        ("arms", (C.c_void_p * 4) * 2),
        ("wta", C.c_void_p * 2),
        ("lrc", C.c_void_p),
        ("vote", C.c_void_p * 5),
        ("interp", C.c_void_p),
        ("discont", C.c_void_p),
        ("final_disp", C.c_void_p),
        ("t_init", C.c_double),
        ("t_agg", C.c_double),
        ("t_scan", C.c_double),
        ("t_multi", C.c_double),
    ]


@dataclass
class Stages:
    """All stage outputs of one ADCensus run.  Volumes are [H][W][Dn] fp32."""

    H: int
    W: int
    Dn: int
    vol_init: list = field(default_factory=list)
    vol_agg: list = field(default_factory=list)
    vol_scan: list = field(default_factory=list)
    arms: list = field(default_factory=list)  # [view][up,down,left,right] int32 [H][W]
    wta: list = field(default_factory=list)
    lrc: np.ndarray | None = None
    vote: list = field(default_factory=list)
    interp: np.ndarray | None = None
    discont: np.ndarray | None = None
    final: np.ndarray | None = None
    seconds: dict = field(default_factory=dict)


def _alloc_stages(H, W, Dn, volumes, vol_shape):
    st = Stages(H, W, Dn)
    t = _Taps()
    if volumes:
        for name in ("vol_init", "vol_agg", "vol_scan"):
            arrs = [np.empty(vol_shape, np.float32) for _ in range(2)]
            setattr(st, name, arrs)
            for k in range(2):
                getattr(t, name)[k] = _p(arrs[k])
    st.arms = [[np.empty((H, W), np.int32) for _ in range(4)] for _ in range(2)]
    for k in range(2):
        for a in range(4):
            t.arms[k][a] = _p(st.arms[k][a])
    st.wta = [np.empty((H, W), np.int32) for _ in range(2)]
    for k in range(2):
        t.wta[k] = _p(st.wta[k])
    st.lrc = np.empty((H, W), np.int32)
    t.lrc = _p(st.lrc)
    st.vote = [np.empty((H, W), np.int32) for _ in range(5)]
    for i in range(5):
        t.vote[i] = _p(st.vote[i])
    st.interp = np.empty((H, W), np.int32)
    t.interp = _p(st.interp)
    st.discont = np.empty((H, W), np.int32)
    t.discont = _p(st.discont)
    st.final = np.empty((H, W), np.float32)
    t.final_disp = _p(st.final)
    return st, t


def _check_pair(left, right):
    left = np.ascontiguousarray(left, np.uint8)
    right = np.ascontiguousarray(right, np.uint8)
    if left.ndim != 3 or left.shape[2] != 3 or left.shape != right.shape:
        raise ValueError("expected two HxWx3 uint8 images of equal size")
    return left, right


class Port:
    """ctypes view of oracle/liboracle.so."""

    def __init__(self, path: Path = PORT_SO):
        if not path.exists():
            build(ref=False)
        self.lib = C.CDLL(str(path))
        self.lib.orc_adcensus.restype = C.c_int

    def run(self, left, right, max_disp: int, volumes: bool = True) -> Stages:
        left, right = _check_pair(left, right)
        H, W, _ = left.shape
        Dn = max_disp + 1
        st, t = _alloc_stages(H, W, Dn, volumes, (H, W, Dn))
        rc = self.lib.orc_adcensus(_p(left), _p(right), H, W, max_disp, C.byref(t))
        if rc != 0:
            raise RuntimeError(f"orc_adcensus failed rc={rc}")
        st.seconds = dict(init=t.t_init, agg=t.t_agg, scan=t.t_scan, multi=t.t_multi)
        return st

    def compute(self, left, right, max_disp: int) -> np.ndarray:
        return self.run(left, right, max_disp, volumes=False).final

    @property
    def threads(self) -> int:
        return int(self.lib.orc_omp_max_threads())

    def set_threads(self, n: int) -> None:
        self.lib.orc_omp_set_num_threads(int(n))


class Ref:
    """ctypes view of oracle/_ref/libadcensus_ref.so (the unmodified reference)."""

    def __init__(self, path: Path = REF_SO):
        if not path.exists():
            build(port=False)
        if not path.exists():
            raise FileNotFoundError(f"{path} missing and {REFERENCE_ROOT} not available to build it")
        self.lib = C.CDLL(str(path))

    def run(self, left, right, max_disp: int, serial_scanline: bool = True, volumes: bool = True, model: str = "RGB",
            min_disp: int = 0) -> Stages:
        """Staged run with taps; volumes are transposed to [H][W][Dn] for comparison (Dn = max_disp - min_disp + 1 planes).
        model "HSI" runs the reference's HSI preprocessing first (bgr2hsi + computeGaussMedian); st.pre = the two
        preprocessed images."""
        left, right = _check_pair(left, right)
        H, W, _ = left.shape
        Dn = max_disp - min_disp + 1
        st, t = _alloc_stages(H, W, Dn, volumes, (Dn, H, W))
        pre = [np.empty((H, W, 3), np.uint8), np.empty((H, W, 3), np.uint8)]
        rc = self.lib.ref_adcensus_staged_model(_p(left), _p(right), H, W, min_disp, max_disp, int(serial_scanline),
                                                {"RGB": 0, "HSI": 1}[model], _p(pre[0]), _p(pre[1]), C.byref(t))
        st.pre = pre
        if rc != 0:
            raise RuntimeError(f"ref_adcensus_staged failed rc={rc}")
        if volumes:
            for name in ("vol_init", "vol_agg", "vol_scan"):
                setattr(st, name, [np.ascontiguousarray(v.transpose(1, 2, 0)) for v in getattr(st, name)])
        st.seconds = dict(init=t.t_init, agg=t.t_agg, scan=t.t_scan, multi=t.t_multi)
        return st

    def compute(self, left, right, max_disp: int):
        """The reference's public ADCensus::compute, as shipped (all threads, racy scanline)."""
        left, right = _check_pair(left, right)
        H, W, _ = left.shape
        out = np.empty((H, W), np.float32)
        sec = C.c_double(0)
        rc = self.lib.ref_adcensus_compute(_p(left), _p(right), H, W, 0, max_disp, _p(out), C.byref(sec))
        if rc != 0:
            raise RuntimeError(f"ref_adcensus_compute failed rc={rc}")
        return out, sec.value

    def compute_ex(self, left, right, max_disp: int, model: str = "RGB", roi: bool = False, mask: bool = False, offset: int = 0,
                   serial: bool = True):
        """ADCensus::compute with the full matching strategy; serial = one OpenMP thread (deterministic)."""
        left, right = _check_pair(left, right)
        H, W, _ = left.shape
        out = np.empty((H, W), np.float32)
        rc = self.lib.ref_adcensus_compute_ex(_p(left), _p(right), H, W, 0, max_disp, {"RGB": 0, "HSI": 1}[model], int(roi), int(mask),
                                              int(offset), int(serial), _p(out))
        if rc != 0:
            raise RuntimeError(f"ref_adcensus_compute_ex failed rc={rc}")
        return out

    def ad_census_pairs(self, left, right, y, xl, xr):
        left, right = _check_pair(left, right)
        H, W, _ = left.shape
        y = np.ascontiguousarray(y, np.int32)
        xl = np.ascontiguousarray(xl, np.int32)
        xr = np.ascontiguousarray(xr, np.int32)
        n = len(y)
        ad3 = np.empty(n, np.int32)
        cen = np.empty(n, np.int32)
        cost = np.empty(n, np.float32)
        self.lib.ref_ad_census_pairs(_p(left), _p(right), H, W, n, _p(y), _p(xl), _p(xr), _p(ad3), _p(cen), _p(cost))
        return ad3, cen, cost

    @property
    def threads(self) -> int:
        return int(self.lib.ref_omp_max_threads())

    def set_threads(self, n: int) -> None:
        """OpenMP threads of the following runs (torchrun exports OMP_NUM_THREADS=1 to its workers)."""
        self.lib.ref_omp_set_num_threads(int(n))


def have_ref() -> bool:
    return REF_SO.exists() or (REFERENCE_ROOT / "source" / "ADCensus.cpp").exists()
