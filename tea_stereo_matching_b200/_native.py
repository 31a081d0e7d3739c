"""ctypes binding of include/tsm.h (libtsm_b200.so).

The shared library is built in-tree by ``tea_stereo_matching_b200.build_native()``
(nvcc, sm_100a only).  There is no CPU fallback: if the library is missing, or no
CUDA device is present, every compute entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("TSM_LIB", PKG_DIR / "libtsm_b200.so"))  # TSM_LIB: kernel experiments only
CSRC_DIR = PKG_DIR / "csrc"

TSM_OK, TSM_E_ARG, TSM_E_CUDA, TSM_E_OOM, TSM_E_UNSUPPORTED, TSM_E_STATE = range(6)

# enum tsm_stage
STAGE_PREP, STAGE_INIT, STAGE_AGGREGATE, STAGE_SCANLINE, STAGE_WTA = 1, 2, 4, 8, 16
STAGE_LRC, STAGE_VOTE, STAGE_INTERP, STAGE_DISCONT, STAGE_SUBPIXEL = 32, 64, 128, 256, 512
STAGE_ALL = 1023
# enum tsm_buffer
(BUF_VOL_LEFT, BUF_VOL_RIGHT, BUF_ARMS_LEFT, BUF_ARMS_RIGHT, BUF_WTA_LEFT, BUF_WTA_RIGHT, BUF_DISP, BUF_EDGES,
 BUF_FINAL, BUF_CENSUS_LEFT, BUF_CENSUS_RIGHT, BUF_IMG_LEFT, BUF_IMG_RIGHT, BUF_IMG4_LEFT, BUF_IMG4_RIGHT) = range(15)
MAP_FIXED, MAP_FLOAT = 0, 1

# every symbol include/tsm.h declares (tests check the .so exports them all)
EXPORTS = [
    "tsm_version", "tsm_status_string", "tsm_last_error", "tsm_create", "tsm_create_on_stream", "tsm_destroy",
    "tsm_device_count", "tsm_synchronize", "tsm_adcensus_compute", "tsm_adcensus_compute_device",
    "tsm_adcensus_enqueue", "tsm_adcensus_wait", "tsm_remap", "tsm_rectify_stereo", "tsm_rectify_adcensus",
    "tsm_rectify_adcensus_device", "tsm_rectify_adcensus_enqueue",
    "tsm_invalidate_maps", "tsm_stage_begin", "tsm_stage_run", "tsm_volume_pitch", "tsm_buffer_bytes", "tsm_tap",
    "tsm_poke", "tsm_set_profiling", "tsm_get_stage_times", "tsm_launch_count", "tsm_selftest",
    "tsm_init_undistort_rectify_map", "tsm_reproject_to_depth", "tsm_reproject_to_3d", "tsm_reproject_to_3d_q", "tsm_apply_colormap", "tsm_jet_colormap", "tsm_write_point_cloud",
]


class Config(C.Structure):
    """struct tsm_adcensus_config"""

    _fields_ = [
        ("min_disparity", C.c_int32),
        ("max_disparity", C.c_int32),
        ("color_model", C.c_int32),
        ("roi_matching", C.c_int32),
        ("mask_matching", C.c_int32),
        ("offset", C.c_int32),
    ]


class NativeLibraryMissing(RuntimeError):
    pass


def build_native(verbose: bool = False) -> Path:
    """Compile every CUDA source for sm_100a into libtsm_b200.so (make + nvcc)."""
    env = dict(os.environ)
    env.pop("CXX", None)
    env.pop("CC", None)
    r = subprocess.run(["make", "-j8", "-C", str(CSRC_DIR)], env=env, capture_output=not verbose, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libtsm_b200.so failed:\n" + (r.stdout or "") + (r.stderr or ""))
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise NativeLibraryMissing(
            f"{LIB_PATH} is missing: build it with tea_stereo_matching_b200.build_native() "
            "(python -c 'import __graft_entry__ as g; g.build()').  There is no CPU fallback."
        )
    L = C.CDLL(str(LIB_PATH))
    vp, sz, i32, u8p = C.c_void_p, C.c_size_t, C.c_int, C.c_void_p
    cfgp = C.POINTER(Config)
    L.tsm_version.restype = C.c_int
    L.tsm_status_string.restype = C.c_char_p
    L.tsm_status_string.argtypes = [i32]
    L.tsm_last_error.restype = C.c_char_p
    L.tsm_last_error.argtypes = [vp]
    L.tsm_create.argtypes = [i32, C.POINTER(vp)]
    L.tsm_create_on_stream.argtypes = [i32, vp, C.POINTER(vp)]
    L.tsm_destroy.argtypes = [vp]
    L.tsm_destroy.restype = None
    L.tsm_device_count.argtypes = [C.POINTER(i32)]
    L.tsm_synchronize.argtypes = [vp]
    L.tsm_adcensus_compute.argtypes = [vp, cfgp, u8p, sz, u8p, sz, i32, i32, vp, sz]
    L.tsm_adcensus_compute_device.argtypes = [vp, cfgp, vp, vp, i32, i32, vp]
    L.tsm_adcensus_enqueue.argtypes = [vp, cfgp, u8p, sz, u8p, sz, i32, i32]
    L.tsm_adcensus_wait.argtypes = [vp, vp, sz]
    u64 = C.c_ulonglong
    L.tsm_remap.argtypes = [vp, u8p, sz, i32, i32, vp, vp, i32, u64, i32, i32, u8p, sz]
    L.tsm_rectify_stereo.argtypes = [vp, u8p, sz, i32, i32, vp, vp, vp, vp, i32, u64, u8p, sz, u8p, sz]
    L.tsm_rectify_adcensus.argtypes = [vp, cfgp, u8p, sz, i32, i32, vp, vp, vp, vp, i32, u64, vp, sz]
    L.tsm_rectify_adcensus_enqueue.argtypes = [vp, cfgp, u8p, sz, i32, i32, vp, vp, vp, vp, i32, u64]
    L.tsm_rectify_adcensus_device.argtypes = [vp, cfgp, vp, sz, i32, i32, vp, vp, vp, vp, i32, u64, vp]
    L.tsm_invalidate_maps.argtypes = [vp]
    L.tsm_invalidate_maps.restype = None
    L.tsm_stage_begin.argtypes = [vp, cfgp, u8p, sz, u8p, sz, i32, i32]
    L.tsm_stage_run.argtypes = [vp, i32, i32]
    L.tsm_volume_pitch.argtypes = [vp]
    L.tsm_buffer_bytes.argtypes = [vp, i32]
    L.tsm_buffer_bytes.restype = sz
    L.tsm_tap.argtypes = [vp, i32, vp, sz]
    L.tsm_poke.argtypes = [vp, i32, vp, sz]
    L.tsm_set_profiling.argtypes = [vp, i32]
    L.tsm_get_stage_times.argtypes = [vp, C.POINTER(i32), C.POINTER(C.c_char_p), C.POINTER(C.c_float)]
    L.tsm_launch_count.argtypes = [vp]
    L.tsm_launch_count.restype = C.c_longlong
    L.tsm_selftest.argtypes = [vp, i32, C.POINTER(C.c_ulonglong)]
    f32 = C.c_float
    L.tsm_reproject_to_depth.argtypes = [vp, vp, sz, i32, i32, f32, f32, vp, sz]
    L.tsm_reproject_to_3d.argtypes = [vp, vp, sz, i32, i32, f32, f32, f32, f32, vp, sz]
    L.tsm_reproject_to_3d_q.argtypes = [vp, vp, sz, i32, i32, C.POINTER(C.c_double), vp, sz]
    L.tsm_apply_colormap.argtypes = [vp, vp, sz, i32, i32, i32, f32, f32, vp, vp, sz]
    dp = C.POINTER(C.c_double)
    L.tsm_init_undistort_rectify_map.argtypes = [vp, dp, dp, i32, dp, dp, i32, i32, i32, vp, sz, vp, sz]
    L.tsm_write_point_cloud.argtypes = [vp, sz, vp, sz, i32, i32, C.c_char_p, i32, C.POINTER(sz)]
    L.tsm_jet_colormap.argtypes = [vp]
    L.tsm_jet_colormap.restype = None
    _lib = L
    return L
