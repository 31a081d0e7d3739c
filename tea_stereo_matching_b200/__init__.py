"""tea_stereo_matching_b200 -- B200-native (sm_100a) AD-Census stereo disparity path.

Host-side mirror of the reference's operator interface for this one path
(stereo::ADCensus, stereo::EpipolarRectify) on top of the C-ABI in include/tsm.h
(libtsm_b200.so, hand-written CUDA kernels under csrc/).  No CPU fallback.
"""
from ._native import LIB_PATH, NativeLibraryMissing, build_native, lib
from .adcensus import ADCensus, ADCensusError, ColorModel, Context, StageRunner
from .consumers import (JETColorMap, applyColorMap, reprojectTo3D, reprojectToDepth, writePointCloudToPCD,
                        writePointCloudToPLY)
from .rectify import (CameraIntrinsic, EpipolarRectify, EpipolarRectifyMap, StereoPair, StereoParams,
                      initUndistortRectifyMap)

__all__ = [
    "ADCensus", "ADCensusError", "ColorModel", "Context", "StageRunner", "EpipolarRectify", "EpipolarRectifyMap",
    "CameraIntrinsic", "StereoPair", "StereoParams", "initUndistortRectifyMap", "JETColorMap", "applyColorMap", "reprojectToDepth", "reprojectTo3D", "writePointCloudToPCD", "writePointCloudToPLY",
    "build_native", "lib", "LIB_PATH", "NativeLibraryMissing",
]
