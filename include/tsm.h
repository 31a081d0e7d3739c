/*
 * include/tsm.h -- C-ABI of the B200-native ADCensus / EpipolarRectify path.
 *
 * This is the drop-in boundary: plain pointers, sizes and int error codes, no
 * C++ or torch types.  The C++20 facade (tea_stereo_matching_b200/cpp/stereo.h)
 * keeps the reference's class signatures and calls only these entry points;
 * INTEGRATION.md shows the binding a maintainer of the reference would add.
 * Every entry point cites the reference interface it replaces (paths relative
 * to the reference repository root).
 *
 * Threading: a tsm_ctx is bound to one CUDA device and one stream and is not
 * thread-safe; use one ctx per worker thread per GPU (the reference's
 * ADCensus object is likewise not re-entrant: mutable ADCensusImpl,
 * source/ADCensus.cpp:275-296).  All functions return TSM_OK (0) or a
 * TSM_E_* code; tsm_last_error() gives the message.  There is NO CPU
 * fallback anywhere behind this header: without a CUDA device every call
 * except tsm_version / tsm_last_error / tsm_status_string fails with TSM_E_CUDA.
 */
#ifndef TSM_H
#define TSM_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define TSM_VERSION 100 /* 0.1.0 */

typedef struct tsm_ctx tsm_ctx;

enum tsm_status {
    TSM_OK = 0,
    TSM_E_ARG = 1,         /* bad argument (NULL, size mismatch, empty image, min>=max, ...) */
    TSM_E_CUDA = 2,        /* CUDA runtime / launch failure, or no device */
    TSM_E_OOM = 3,         /* device or pinned-host allocation failed */
    TSM_E_UNSUPPORTED = 4, /* accepted by the reference's setters but outside what is built: more than 768 cost planes (ROI / mask
                              matching on images wider than 1535 px), negative disparities, max < 2 * min (the reference's own
                              WTA range is empty there, ADCensus.cpp:1398) */
    TSM_E_STATE = 5        /* call sequence error (wait without enqueue, tap before run, ...) */
};

/* stereo::ColorModel, include/stereo_utils.h:191-195 */
enum tsm_color_model { TSM_COLOR_RGB = 0, TSM_COLOR_HSI = 1 };

/* The state ADCensus keeps between setters and compute():
 *   setMinMaxDisparity  source/ADCensus.cpp:307-313
 *   setMatchingStrategy source/ADCensus.cpp:315-321
 *   setOffset           source/ADCensus.cpp:323-328
 * Tunables (lambda, tau, L, pi, voting, Canny) are the RGB constants of
 * ADCensusParams::setADCensusParams, source/stereo_utils.cpp:271-326. */
typedef struct tsm_adcensus_config {
    int32_t min_disparity; /* >= 0; with min > 0 the reference's plane-index / disparity mix is reproduced as it is (tsm_common.cuh, Dims) */
    int32_t max_disparity; /* Dn = max - min + 1 cost planes (ADCensus.cpp:345) */
    int32_t color_model;   /* enum tsm_color_model; the reference default-constructs HSI (ADCensus.cpp:409-420) */
    int32_t roi_matching;  /* != 0: ROI mode -- max_disparity is replaced by W / 2 (ADCensus.cpp:339-340), HSI images are hue-filtered
                              instead of Gauss-median filtered (:354-360), offset is added and black left pixels are marked -1 (:388-403) */
    int32_t mask_matching; /* != 0: mask mode -- black pixels (0,0,0) are holes: cost 2 (ADCensus.cpp:551-555), census term dropped
                              (:459, 481), zero / clipped arms (:625, 673), skipped scanline steps (:824, 862); plus all of the ROI mode */
    int32_t offset;        /* setOffset: added to every positive disparity in ROI mode (disparityOffset, :1415-1427) */
} tsm_adcensus_config;

/* cv::Mat depth/channel codes of the rectify maps (EpipolarRectifyMap,
 * include/stereo_utils.h:109-148): cv::initUndistortRectifyMap(..., CV_16SC2, ...)
 * emits (CV_16SC2, CV_16UC1) (source/stereo_utils.cpp:164-167); YAML-loaded maps
 * may be (CV_32FC1, CV_32FC1). */
enum tsm_map_kind {
    TSM_MAP_FIXED_16SC2_16UC1 = 0, /* map1 int16[H][W][2] (x,y), map2 uint16[H][W] */
    TSM_MAP_FLOAT_32FC1_X2 = 1     /* map1 float[H][W] (x), map2 float[H][W] (y) */
};

int tsm_version(void);
const char* tsm_status_string(int status);
/* Message of the last failure on this ctx (ctx == NULL: last failure of tsm_create). */
const char* tsm_last_error(const tsm_ctx* ctx);

/* Lifetime.  tsm_create makes its own non-blocking stream on `device`.
 * tsm_create_on_stream borrows the caller's cudaStream_t (e.g. the current
 * torch stream) so the caller can time the kernels with its own events.
 * Replaces ADCensus::ADCensus()/~ADCensus() (source/ADCensus.cpp:298-305). */
int tsm_create(int device, tsm_ctx** out);
int tsm_create_on_stream(int device, void* cuda_stream, tsm_ctx** out);
void tsm_destroy(tsm_ctx* ctx);
int tsm_device_count(int* count);
int tsm_synchronize(tsm_ctx* ctx);

/* ---- stereo::ADCensus::compute(left, right, disparity), source/ADCensus.cpp:330-407 ----
 * left/right: host, CV_8UC3 BGR, H rows of W pixels, row strides lstep/rstep bytes.
 * disparity : host, CV_32FC1, row stride dstep bytes.  Blocking: returns after the
 * result has been copied back.  Valid pixels hold sub-pixel disparities in
 * [min,max]; invalid ones are negative (-1 occlusion / -2 mismatch, median-mixed). */
int tsm_adcensus_compute(tsm_ctx* ctx, const tsm_adcensus_config* cfg,
                         const uint8_t* left, size_t lstep, const uint8_t* right, size_t rstep,
                         int H, int W, float* disparity, size_t dstep);

/* Same computation on device-resident inputs (packed BGR [H][W][3], packed float
 * output), enqueued on the ctx stream without any host synchronisation.  This is
 * the HBM-resident form bench.py times as `value`. */
int tsm_adcensus_compute_device(tsm_ctx* ctx, const tsm_adcensus_config* cfg,
                                const uint8_t* d_left, const uint8_t* d_right, int H, int W, float* d_disparity);

/* Asynchronous pair for batches (frames are sharded over GPUs and
 * several frames are in flight per GPU, one ctx each): enqueue copies the pair
 * into pinned staging, H2D, kernels, D2H into pinned staging; wait blocks and
 * copies the result into the caller's buffer. */
int tsm_adcensus_enqueue(tsm_ctx* ctx, const tsm_adcensus_config* cfg,
                         const uint8_t* left, size_t lstep, const uint8_t* right, size_t rstep, int H, int W);
int tsm_adcensus_wait(tsm_ctx* ctx, float* disparity, size_t dstep);

/* Rectify maps are uploaded once and reused from device memory.  Host addresses cannot identify map CONTENT (a
 * reloaded map may land on the same addresses), so every entry point that takes maps also takes `map_generation`:
 * an id the caller changes whenever the content behind the pointers changes (the facades bump a process-wide
 * counter in EpipolarRectify::loadEpipolarRectifyMap, source/EpipolarRectify.cpp:32-44).  The device copy is reused
 * only while generation, pointers, kind and size are all unchanged; map_generation == 0 uploads on every call. */

/* ---- stereo::EpipolarRectify::rectify(left, right, L, R), source/EpipolarRectify.cpp:87-101 ----
 * One cv::remap(src, dst, map1, map2, INTER_LINEAR) (BORDER_CONSTANT 0) of a CV_8UC3
 * image: src is sH x sW (row stride sstep), dst is H x W (the map size, row stride dstep). */
int tsm_remap(tsm_ctx* ctx, const uint8_t* src, size_t sstep, int sH, int sW,
              const void* map1, const void* map2, int map_kind, unsigned long long map_generation,
              int H, int W, uint8_t* dst, size_t dstep);

/* ---- stereo::EpipolarRectify::rectify(stereoImage, L, R), source/EpipolarRectify.cpp:68-85 ----
 * Crops the side-by-side frame (H x 2W) into halves [0,W) / [W,2W) and remaps each with its
 * own map pair (map00,map01) / (map10,map11). */
int tsm_rectify_stereo(tsm_ctx* ctx, const uint8_t* stereo, size_t sstep, int H, int W,
                       const void* map00, const void* map01, const void* map10, const void* map11, int map_kind,
                       unsigned long long map_generation, uint8_t* left, size_t lstep, uint8_t* right, size_t rstep);

/* Fused rectify -> ADCensus (BASELINE config C4): the rectified pair never leaves
 * the device.  tsm_invalidate_maps() drops the device copies of the maps. */
int tsm_rectify_adcensus(tsm_ctx* ctx, const tsm_adcensus_config* cfg,
                         const uint8_t* stereo, size_t sstep, int H, int W,
                         const void* map00, const void* map01, const void* map10, const void* map11, int map_kind,
                         unsigned long long map_generation, float* disparity, size_t dstep);
/* Asynchronous form for batches: enqueue, then tsm_adcensus_wait() on the same ctx. */
int tsm_rectify_adcensus_enqueue(tsm_ctx* ctx, const tsm_adcensus_config* cfg,
                                 const uint8_t* stereo, size_t sstep, int H, int W,
                                 const void* map00, const void* map01, const void* map10, const void* map11, int map_kind,
                                 unsigned long long map_generation);
/* The same on a device-resident side-by-side frame (row stride sstep bytes) and a packed device output, enqueued
 * on the ctx stream without host synchronisation (the maps stay host pointers: they are uploaded once per
 * generation).  This is the HBM-resident form bench.py times for config C4. */
int tsm_rectify_adcensus_device(tsm_ctx* ctx, const tsm_adcensus_config* cfg,
                                const uint8_t* d_stereo, size_t sstep, int H, int W,
                                const void* map00, const void* map01, const void* map10, const void* map11, int map_kind,
                                unsigned long long map_generation, float* d_disparity);
void tsm_invalidate_maps(tsm_ctx* ctx);

/* ---- stereo::EpipolarRectifyMap::compute, source/stereo_utils.cpp:157-169 ----
 * One cv::initUndistortRectifyMap(cameraMatrix, distCoeffs, R, newCameraMatrix, Size(W, H), CV_16SC2,
 * map1, map2): map1 = CV_16SC2 (H x W x 2 int16), map2 = CV_16UC1, exactly the pair tsm_remap takes as
 * TSM_MAP_FIXED_16SC2_16UC1.  camera_matrix, R, new_camera_matrix: 3x3 row-major doubles (R NULL =
 * identity; new_camera_matrix NULL = camera_matrix with the principal point moved to the image centre, as
 * cv::getDefaultNewCameraMatrix(K, size, true) does; for a 3x4 projection matrix P pass ld_new = 4).
 * dist_coeffs: 0, 4, 5, 8, 12 or 14 doubles in OpenCV order (k1 k2 p1 p2 k3 k4 k5 k6 s1 s2 s3 s4 tx ty). */
int tsm_init_undistort_rectify_map(tsm_ctx* ctx, const double* camera_matrix, const double* dist_coeffs, int n_dist,
                                   const double* R, const double* new_camera_matrix, int ld_new, int H, int W,
                                   int16_t* map1, size_t step1, uint16_t* map2, size_t step2);

/* ---- consumers of the disparity map: the free functions of source/stereo.cpp that follow
 * ADCensus::compute.  `disparity` is a host CV_32FC1 map (row stride `step`
 * bytes), or NULL: then the map the last tsm_adcensus_* call of THIS context produced is consumed
 * where it lies in device memory (H, W must match it) -- no D2H / H2D round trip.  Outputs are host
 * buffers.  Pixels with d < 0 or d == inf stay 0 like in the reference.
 *   tsm_reproject_to_depth   stereo::reprojectToDepth(disparity, f, B, depth)            stereo.cpp:139-151
 *   tsm_reproject_to_3d      stereo::reprojectTo3D(disparity, f, B, cx, cy, XYZ)         stereo.cpp:153-172
 *   tsm_reproject_to_3d_q    stereo::reprojectTo3D(disparity, Q, XYZ), Q = 4x4 row-major stereo.cpp:174-202
 *   tsm_apply_colormap       stereo::applyColorMap(src, dst[, minVal, maxVal], colorMap)  stereo.cpp:95-137
 *                            colormap = 256 x BGR bytes, NULL = stereo::JETColorMap()
 *   tsm_jet_colormap         stereo::JETColorMap()                                        stereo.cpp:75-93
 * xyz is CV_32FC3 (3 floats per pixel), dst of the colour map CV_8UC3. */
int tsm_reproject_to_depth(tsm_ctx* ctx, const float* disparity, size_t step, int H, int W, float focal_length, float baseline,
                           float* depth, size_t dstep);
int tsm_reproject_to_3d(tsm_ctx* ctx, const float* disparity, size_t step, int H, int W, float focal_length, float baseline,
                        float cx, float cy, float* xyz, size_t xstep);
int tsm_reproject_to_3d_q(tsm_ctx* ctx, const float* disparity, size_t step, int H, int W, const double* Q, float* xyz,
                          size_t xstep);
int tsm_apply_colormap(tsm_ctx* ctx, const float* disparity, size_t step, int H, int W, int auto_range, float min_val,
                       float max_val, const uint8_t* colormap, uint8_t* dst, size_t dstep);
void tsm_jet_colormap(uint8_t* table768);

/* ---- stereo::writePointCloudToPCD / writePointCloudToPLY, source/stereo.cpp:204-356 ----
 * ASCII point-cloud files of the finite points of an XYZ map (CV_32FC3, row stride xstep bytes) coloured by a CV_8UC3 BGR
 * image (row stride cstep bytes): points with an infinite coordinate are dropped (:263-266), numbers are written with
 * std::to_chars (shortest round-trip form), PCD packs the colour as r << 16 | g << 8 | b | 1 << 24 (:240).  Host file I/O:
 * no context, no device work.  format: 0 = PCD v0.7, 1 = PLY.  *points (optional) receives the number of points written. */
enum tsm_cloud_format { TSM_CLOUD_PCD = 0, TSM_CLOUD_PLY = 1 };
int tsm_write_point_cloud(const uint8_t* bgr, size_t cstep, const float* xyz, size_t xstep, int H, int W, const char* path,
                          int format, size_t* points);

/* ---- parity taps: the analogue of the reference's writeProcess debug dumps
 * (source/ADCensus.cpp:573-580, 785-792, 1003-1010).  tsm_stage_begin uploads a pair and
 * sizes the arena; tsm_stage_run runs the stages in `mask` (TSM_STAGE_* bits, in pipeline
 * order); tsm_tap / tsm_poke copy an intermediate buffer device->host / host->device so
 * each kernel can be checked on the ORACLE's input for that stage. */
enum tsm_stage {
    TSM_STAGE_PREP = 1 << 0,     /* census signatures, arms, window sizes, similarity flags */
    TSM_STAGE_INIT = 1 << 1,     /* costInitialize       ADCensus.cpp:522-581 */
    TSM_STAGE_AGGREGATE = 1 << 2,/* costAggregate        ADCensus.cpp:753-793 */
    TSM_STAGE_SCANLINE = 1 << 3, /* scanlineOptimize     ADCensus.cpp:997-1011 */
    TSM_STAGE_WTA = 1 << 4,      /* cost2disparity x2    ADCensus.cpp:1394-1413 */
    TSM_STAGE_LRC = 1 << 5,      /* outlierElimination   ADCensus.cpp:1013-1044 */
    TSM_STAGE_VOTE = 1 << 6,     /* regionVoting x5      ADCensus.cpp:1046-1159 (arg = one call 0..4, -1 = all) */
    TSM_STAGE_INTERP = 1 << 7,   /* properInterpolation  ADCensus.cpp:1161-1239 */
    TSM_STAGE_DISCONT = 1 << 8,  /* discontinuityAdjust. ADCensus.cpp:1256-1342 */
    TSM_STAGE_SUBPIXEL = 1 << 9, /* subpixelEnhancement  ADCensus.cpp:1344-1374 */
    TSM_STAGE_ALL = (1 << 10) - 1
};
enum tsm_buffer {
    TSM_BUF_VOL_LEFT = 0,  /* float [H][W][Dn] dense (the device keeps a split, 128-byte aligned layout) */
    TSM_BUF_VOL_RIGHT = 1,
    TSM_BUF_ARMS_LEFT = 2, /* uint8 [H][W][4] = up, down, left, right */
    TSM_BUF_ARMS_RIGHT = 3,
    TSM_BUF_WTA_LEFT = 4,  /* int32 [H][W] */
    TSM_BUF_WTA_RIGHT = 5,
    TSM_BUF_DISP = 6,      /* int32 [H][W], the working disparity map (m_disparityMap) */
    TSM_BUF_EDGES = 7,     /* uint8 [H][W], Canny output of discontinuityAdjustment */
    TSM_BUF_FINAL = 8,     /* float [H][W] */
    TSM_BUF_CENSUS_LEFT = 9,  /* uint64 [6][H][W]: lt_B, lt_G, lt_R, gt_B, gt_G, gt_R */
    TSM_BUF_CENSUS_RIGHT = 10,
    TSM_BUF_IMG_LEFT = 11, /* uint8 [H][W][3] */
    TSM_BUF_IMG_RIGHT = 12,
    TSM_BUF_IMG4_LEFT = 13, /* uint32 [H][W]: the image the matching stages see, channel c in byte c (B,G,R or, for the
                               HSI model, H,S,I after bgr2hsi + computeGaussMedian) */
    TSM_BUF_IMG4_RIGHT = 14
};
int tsm_stage_begin(tsm_ctx* ctx, const tsm_adcensus_config* cfg,
                    const uint8_t* left, size_t lstep, const uint8_t* right, size_t rstep, int H, int W);
int tsm_stage_run(tsm_ctx* ctx, int mask, int arg);
int tsm_volume_pitch(const tsm_ctx* ctx); /* floats per pixel of a tapped volume (= Dn) */
size_t tsm_buffer_bytes(const tsm_ctx* ctx, int buffer);
int tsm_tap(tsm_ctx* ctx, int buffer, void* dst, size_t bytes);
int tsm_poke(tsm_ctx* ctx, int buffer, const void* src, size_t bytes);

/* Device self-tests of arithmetic shortcuts the kernels rely on.  *mismatches receives the number of
 * failing cases (0 = pass).  TSM_SELFTEST_DIV: the branch-free fp32 divide used by the aggregation
 * normalisation equals the IEEE divide (source/ADCensus.cpp:747) for every divisor 1..4489. */
enum tsm_selftest_id { TSM_SELFTEST_DIV = 0 };
int tsm_selftest(tsm_ctx* ctx, int which, unsigned long long* mismatches);

/* Per-stage device time of the last compute on this ctx, in milliseconds, measured
 * with CUDA events on the ctx stream when profiling was enabled.  names/ms arrays of
 * length *n; returns the number of stages actually filled in *n. */
int tsm_set_profiling(tsm_ctx* ctx, int enabled);
int tsm_get_stage_times(tsm_ctx* ctx, int* n, const char** names, float* ms);
/* Number of kernels this library launched on the ctx since creation (bench.py's gpu_launches). */
long long tsm_launch_count(const tsm_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif
