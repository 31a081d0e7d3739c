/*
 * oracle/ref_driver.cpp -- TEST INFRASTRUCTURE (oracle), not product code.
 *
 * Builds the UNMODIFIED reference translation unit
 *   /root/reference/source/ADCensus.cpp
 * (included below from where it lies; nothing is copied into this repo) into
 * oracle/_ref/libadcensus_ref.so, against the cv:: shim in oracle/ref_shim/.
 * It adds only:
 *   - link stubs for the reference's logger / utils / StereoMatching symbols
 *     (their real sources need MSVC's localtime_s: source/utils.cpp:159,170);
 *   - ADCensusParams::setADCensusParams, spliced in at build time from
 *     source/stereo_utils.cpp:271-326 via REF_PARAMS_INC (a temp file outside
 *     the repo, written by oracle/Makefile with sed);
 *   - an extern "C" entry that runs the reference's own stage functions in
 *     the order of ADCensus::compute (ADCensus.cpp:372-391) and
 *     multiOptimize (ADCensus.cpp:1376-1392) and memcpy-taps the results.
 *
 * Determinism: the reference's OpenMP scanline is racy (ADCensus.cpp:801-815,
 * 837-853 parallelise loops with a loop-carried in-place dependency).  Parity
 * runs pass serial_scanline=1, which wraps scanlineOptimize() in
 * omp_set_num_threads(1); timing runs pass 0 (as shipped).
 */
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <limits>
#include <memory>
#include <mutex>
#include <string>
#include <vector>
#include <omp.h>
#include "ref_shim/cvshim.hpp"

#define private public
#include "include/stereo.h" /* resolved with -I$(REF) = /root/reference */
#undef private
#include "include/utils.h"
#include "include/logger.h"

/* ---- link stubs for reference symbols whose real TUs cannot build here ---- */
namespace logging {
std::shared_ptr<Logger> Logger::instance;
std::mutex Logger::createMtx;
Logger::Logger() {}
Logger::~Logger() {}
std::shared_ptr<Logger> Logger::getInstance()
{
    std::lock_guard<std::mutex> lk(createMtx);
    if (!instance) instance = std::shared_ptr<Logger>(new Logger(), [](Logger*) {});
    return instance;
}
void Logger::log(LogLevel, const std::string& msg, const std::string&, const std::string&, int)
{
    if (std::getenv("ORACLE_REF_LOG")) std::cerr << "[ref] " << msg << std::endl;
}
}  // namespace logging
namespace utils {
std::string formatMilliseconds(double ms) { return std::to_string(ms); }
}
stereo::StereoMatching::~StereoMatching() {}
#include REF_PARAMS_INC /* = source/stereo_utils.cpp:271-326, ADCensusParams::setADCensusParams */

/* ---- the reference implementation itself, unmodified ---- */
#include "source/ADCensus.cpp" /* resolved with -I$(REF) = /root/reference */

namespace {
double now_s()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
/* model: 0 = RGB, 1 = HSI (plain: no ROI / mask), the preprocessing of ADCensus::compute, ADCensus.cpp:350-371 */
void set_images(stereo::ADCensus& a, const uint8_t* left, const uint8_t* right, int H, int W, int minD, int maxD, int model = 0)
{
    a.setMatchingStrategy(model == 1 ? stereo::ColorModel::HSI : stereo::ColorModel::RGB, false, false);
    a.setMinMaxDisparity(minD, maxD);
    auto& im = *a.impl;
    cv::Mat L(cv::Size(W, H), CV_8UC3), R(cv::Size(W, H), CV_8UC3);
    std::memcpy(L.data, left, (size_t)H * W * 3);
    std::memcpy(R.data, right, (size_t)H * W * 3);
    if (model == 1) {
        cv::Mat hl, hr, fl, fr;
        im.bgr2hsi(L, hl, false);
        im.bgr2hsi(R, hr, false);
        im.computeGaussMedian(hl, fl, 3);
        im.computeGaussMedian(hr, fr, 3);
        L = fl.clone();
        R = fr.clone();
    }
    im.m_images[0] = L;
    im.m_images[1] = R;
    im.m_imageSize = cv::Size(W, H);
    const int Dn = maxD - minD + 1;
    im.m_costMaps.resize(2);
    for (int k = 0; k < 2; ++k) {
        im.m_costMaps[k].resize(Dn);
        for (int d = 0; d < Dn; ++d) im.m_costMaps[k][d].create(im.m_imageSize, CV_32F);
    }
}
void tap_volume(stereo::ADCensus& a, int view, float* dst, int H, int W, int Dn)
{
    if (!dst) return;
    for (int d = 0; d < Dn; ++d)
        std::memcpy(dst + (size_t)d * H * W, a.impl->m_costMaps[view][d].data, (size_t)H * W * sizeof(float));
}
void tap_i32(const cv::Mat& m, int32_t* dst)
{
    if (dst) std::memcpy(dst, m.data, (size_t)m.rows * m.cols * sizeof(int32_t));
}
}  // namespace

extern "C" {

/* All tap pointers are optional (NULL = skip).  Volumes are in the reference's
 * own layout [Dn][H][W] fp32 (ADCensus.cpp:289); maps are [H][W]. */
struct RefTaps {
    float* vol_init[2];
    float* vol_agg[2];
    float* vol_scan[2];
    int32_t* arms[2][4]; /* [view][up,down,left,right] */
    int32_t* wta[2];
    int32_t* lrc;
    int32_t* vote[5];
    int32_t* interp;
    int32_t* discont;
    float* final_disp;
    double t_init, t_agg, t_scan, t_multi; /* seconds, out */
};

int ref_adcensus_staged_model(const uint8_t* left, const uint8_t* right, int H, int W, int minD, int maxD,
                              int serial_scanline, int model, uint8_t* pre_left, uint8_t* pre_right, RefTaps* t);

int ref_adcensus_staged(const uint8_t* left, const uint8_t* right, int H, int W, int minD, int maxD,
                        int serial_scanline, RefTaps* t)
{
    return ref_adcensus_staged_model(left, right, H, W, minD, maxD, serial_scanline, 0, nullptr, nullptr, t);
}

/* Same with the colour model selectable (0 RGB, 1 HSI); pre_left / pre_right (optional, H*W*3 bytes) receive the
 * images the matching stages actually see (HSI: bgr2hsi + computeGaussMedian). */
int ref_adcensus_staged_model(const uint8_t* left, const uint8_t* right, int H, int W, int minD, int maxD,
                              int serial_scanline, int model, uint8_t* pre_left, uint8_t* pre_right, RefTaps* t)
{
    try {
        stereo::ADCensus a;
        set_images(a, left, right, H, W, minD, maxD, model);
        auto& im = *a.impl;
        if (pre_left) std::memcpy(pre_left, im.m_images[0].data, (size_t)H * W * 3);
        if (pre_right) std::memcpy(pre_right, im.m_images[1].data, (size_t)H * W * 3);
        const int Dn = maxD - minD + 1;
        double t0 = now_s();
        im.costInitialize();
        double t1 = now_s();
        tap_volume(a, 0, t->vol_init[0], H, W, Dn);
        tap_volume(a, 1, t->vol_init[1], H, W, Dn);
        double t2 = now_s();
        im.costAggregate();
        double t3 = now_s();
        tap_volume(a, 0, t->vol_agg[0], H, W, Dn);
        tap_volume(a, 1, t->vol_agg[1], H, W, Dn);
        for (int k = 0; k < 2; ++k) {
            tap_i32(im.m_upLimits[k], t->arms[k][0]);
            tap_i32(im.m_downLimits[k], t->arms[k][1]);
            tap_i32(im.m_leftLimits[k], t->arms[k][2]);
            tap_i32(im.m_rightLimits[k], t->arms[k][3]);
        }
        const int nthreads = omp_get_max_threads();
        double t4 = now_s();
        if (serial_scanline) omp_set_num_threads(1);
        im.scanlineOptimize();
        if (serial_scanline) omp_set_num_threads(nthreads);
        double t5 = now_s();
        tap_volume(a, 0, t->vol_scan[0], H, W, Dn);
        tap_volume(a, 1, t->vol_scan[1], H, W, Dn);
        /* multiOptimize, ADCensus.cpp:1376-1392, step by step */
        double t6 = now_s();
        cv::Mat disp0, disp1;
        im.cost2disparity(0, disp0);
        im.cost2disparity(1, disp1);
        tap_i32(disp0, t->wta[0]);
        tap_i32(disp1, t->wta[1]);
        im.m_disparityMap = im.outlierElimination(disp0, disp1);
        tap_i32(im.m_disparityMap, t->lrc);
        bool hf = false;
        for (int i = 0; i < 5; ++i) {
            im.regionVoting(im.m_disparityMap, im.m_upLimits, im.m_downLimits, im.m_leftLimits, im.m_rightLimits, hf);
            tap_i32(im.m_disparityMap, t->vote[i]);
            hf = !hf;
        }
        im.properInterpolation(im.m_disparityMap, im.m_images[0]);
        tap_i32(im.m_disparityMap, t->interp);
        im.discontinuityAdjustment(im.m_disparityMap, im.m_costMaps);
        tap_i32(im.m_disparityMap, t->discont);
        im.m_floatDisparityMap = im.subpixelEnhancement(im.m_disparityMap, im.m_costMaps);
        double t7 = now_s();
        if (t->final_disp)
            std::memcpy(t->final_disp, im.m_floatDisparityMap.data, (size_t)H * W * sizeof(float));
        t->t_init = t1 - t0; t->t_agg = t3 - t2; t->t_scan = t5 - t4; t->t_multi = t7 - t6;
        return 0;
    } catch (const std::string& e) {
        std::cerr << "[ref] " << e << std::endl;
        return -1;
    } catch (const std::exception& e) {
        std::cerr << "[ref] " << e.what() << std::endl;
        return -2;
    }
}

/* The reference's public entry point, exactly as a user calls it (as shipped:
 * all OpenMP threads, racy scanline).  Returns wall seconds via *seconds. */
int ref_adcensus_compute(const uint8_t* left, const uint8_t* right, int H, int W, int minD, int maxD,
                         float* out, double* seconds)
{
    try {
        stereo::ADCensus a;
        a.setMatchingStrategy(stereo::ColorModel::RGB, false, false);
        a.setMinMaxDisparity(minD, maxD);
        cv::Mat L(cv::Size(W, H), CV_8UC3), R(cv::Size(W, H), CV_8UC3), D;
        std::memcpy(L.data, left, (size_t)H * W * 3);
        std::memcpy(R.data, right, (size_t)H * W * 3);
        double t0 = now_s();
        a.compute(L, R, D);
        double t1 = now_s();
        if (seconds) *seconds = t1 - t0;
        if (out) std::memcpy(out, D.data, (size_t)H * W * sizeof(float));
        return 0;
    } catch (const std::string& e) {
        std::cerr << "[ref] " << e << std::endl;
        return -1;
    } catch (const std::exception& e) {
        std::cerr << "[ref] " << e.what() << std::endl;
        return -2;
    }
}

/* The public entry point with the full matching strategy (model 0 RGB / 1 HSI, roi, mask, offset); serial != 0 runs it
 * with one OpenMP thread (deterministic scanline).  maxD is what setMinMaxDisparity gets; ROI / mask modes replace it by
 * W / 2 inside compute (ADCensus.cpp:339-340). */
int ref_adcensus_compute_ex(const uint8_t* left, const uint8_t* right, int H, int W, int minD, int maxD, int model, int roi,
                            int mask, int offset, int serial, float* out)
{
    const int nthreads = omp_get_max_threads();
    try {
        stereo::ADCensus a;
        a.setMatchingStrategy(model == 1 ? stereo::ColorModel::HSI : stereo::ColorModel::RGB, roi != 0, mask != 0);
        a.setMinMaxDisparity(minD, maxD);
        a.setOffset(offset);
        cv::Mat L(cv::Size(W, H), CV_8UC3), R(cv::Size(W, H), CV_8UC3), D;
        std::memcpy(L.data, left, (size_t)H * W * 3);
        std::memcpy(R.data, right, (size_t)H * W * 3);
        if (serial) omp_set_num_threads(1);
        a.compute(L, R, D);
        if (serial) omp_set_num_threads(nthreads);
        if (out) std::memcpy(out, D.data, (size_t)H * W * sizeof(float));
        return 0;
    } catch (const std::string& e) {
        omp_set_num_threads(nthreads);
        std::cerr << "[ref] " << e << std::endl;
        return -1;
    } catch (const std::exception& e) {
        omp_set_num_threads(nthreads);
        std::cerr << "[ref] " << e.what() << std::endl;
        return -2;
    }
}

/* Integer AD sums and census counts for n (y, xL, xR) pairs, straight from the
 * reference's computeRGBADCost / computeRGBCensusCost (ADCensus.cpp:426,454).
 * ad3 = round(ad * 3) is the integer |dB|+|dG|+|dR|. */
int ref_ad_census_pairs(const uint8_t* left, const uint8_t* right, int H, int W, int n,
                        const int32_t* y, const int32_t* xl, const int32_t* xr,
                        int32_t* ad3, int32_t* census, float* cost)
{
    stereo::ADCensus a;
    set_images(a, left, right, H, W, 0, 1);
    for (int i = 0; i < n; ++i) {
        float ad = a.impl->computeRGBADCost(y[i], xl[i], y[i], xr[i]);
        float ce = a.impl->computeRGBCensusCost(y[i], xl[i], y[i], xr[i], 7, 9);
        ad3[i] = (int32_t)std::lround(ad * 3.0);
        census[i] = (int32_t)ce;
        cost[i] = a.impl->computeADCensusCost(y[i], xl[i], y[i], xr[i], 7, 9);
    }
    return 0;
}

int ref_omp_max_threads(void) { return omp_get_max_threads(); }
/* Launchers such as torchrun export OMP_NUM_THREADS=1; a timing run sets the thread count it reports explicitly. */
void ref_omp_set_num_threads(int n) { if (n > 0) omp_set_num_threads(n); }

}  // extern "C"
