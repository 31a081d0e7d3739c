// k_rectify.cu -- cv::remap(src, dst, map1, map2, INTER_LINEAR, BORDER_CONSTANT 0) for
// CV_8UC3, the one OpenCV call behind stereo::EpipolarRectify::rectify (reference
// source/EpipolarRectify.cpp:99-100).  Fixed-point form of OpenCV 4.13's remapBilinear:
//   (sx, sy) = map1 (CV_16SC2), fx = map2 & 31, fy = (map2 >> 5) & 31
//   dst = (sum_{4 taps} a_i * b_j * src + 512) >> 10,  a = {32-fx, fx}, b = {32-fy, fy}
// out-of-image taps read 0.  Float maps are first quantised with cvRound(v * 32)
// (convertMaps), which is also what cv::remap does internally.
//
// Mapping: one CTA per 32x8 output tile.  The source footprint of a rectification tile
// is a slightly rotated / distorted copy of the tile, so its bounding box is staged in
// shared memory (BGR bytes, coalesced row reads) and the four bilinear taps are gathered
// from there; tiles whose footprint does not fit fall back to direct global gathers.
#include "tsm_common.cuh"
#include <limits.h>

namespace tsm {

constexpr int RM_TW = 32, RM_TH = 8;
constexpr int RM_SW = 64, RM_SH = 24;  // staged source window (pixels)

__global__ void __launch_bounds__(RM_TW* RM_TH)
k_remap(const uint8_t* __restrict__ src, size_t sstep, int sH, int sW, const short2* __restrict__ map1,
        const uint16_t* __restrict__ map2, int H, int W, uint8_t* __restrict__ dst, size_t dstep)
{
    __shared__ uint8_t win[RM_SH][RM_SW * 3];
    __shared__ int bb[4];  // xmin, ymin, xmax, ymax of the tile's taps
    const int tid = threadIdx.y * RM_TW + threadIdx.x;
    const int x = blockIdx.x * RM_TW + threadIdx.x, y = blockIdx.y * RM_TH + threadIdx.y;
    const bool in = x < W && y < H;
    if (tid == 0) { bb[0] = INT_MAX; bb[1] = INT_MAX; bb[2] = INT_MIN; bb[3] = INT_MIN; }
    __syncthreads();
    short2 m1 = make_short2(0, 0);
    int fx = 0, fy = 0;
    if (in) {
        const size_t k = (size_t)y * W + x;
        m1 = map1[k];
        const uint16_t m2 = map2[k];
        fx = m2 & 31;
        fy = (m2 >> 5) & 31;
        atomicMin(&bb[0], (int)m1.x);
        atomicMin(&bb[1], (int)m1.y);
        atomicMax(&bb[2], (int)m1.x + 1);
        atomicMax(&bb[3], (int)m1.y + 1);
    }
    __syncthreads();
    const int bx0 = bb[0], by0 = bb[1];
    const bool staged = (bb[2] - bx0 < RM_SW) && (bb[3] - by0 < RM_SH) && bb[2] >= bx0;
    if (staged) {
        // stage rows by0..by0+RM_SH-1, bytes 3*bx0 .. 3*(bx0+RM_SW)-1, zero outside the image
        for (int i = tid; i < RM_SH * RM_SW * 3; i += RM_TW * RM_TH) {
            const int r = i / (RM_SW * 3), c = i % (RM_SW * 3);
            const int yy = by0 + r, xx3 = 3 * bx0 + c;
            uint8_t v = 0;
            if (yy >= 0 && yy < sH && xx3 >= 0 && xx3 < 3 * sW && yy <= bb[3] && xx3 < 3 * (bb[2] + 1))
                v = src[(size_t)yy * sstep + xx3];
            win[r][c] = v;
        }
    }
    __syncthreads();
    if (!in) return;
    const int wx0 = 32 - fx, wx1 = fx, wy0 = 32 - fy, wy1 = fy;
    const int sx = m1.x, sy = m1.y;
    int acc[3] = {0, 0, 0};
    if (staged) {
        const int lx = sx - bx0, ly = sy - by0;
        const uint8_t* r0 = &win[ly][lx * 3];
        const uint8_t* r1 = &win[ly + 1][lx * 3];
#pragma unroll
        for (int c = 0; c < 3; ++c)
            acc[c] = wx0 * wy0 * r0[c] + wx1 * wy0 * r0[3 + c] + wx0 * wy1 * r1[c] + wx1 * wy1 * r1[3 + c];
    } else {
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int xx = sx + i, yy = sy + j;
                if (xx >= 0 && xx < sW && yy >= 0 && yy < sH) {
                    const uint8_t* s = src + (size_t)yy * sstep + 3 * xx;
                    const int w = (i ? wx1 : wx0) * (j ? wy1 : wy0);
#pragma unroll
                    for (int c = 0; c < 3; ++c) acc[c] += w * s[c];
                }
            }
    }
    uint8_t* o = dst + (size_t)y * dstep + 3 * x;
#pragma unroll
    for (int c = 0; c < 3; ++c) o[c] = (uint8_t)((acc[c] + 512) >> 10);
}

// cv::convertMaps CV_32FC1 x2 -> CV_16SC2 + CV_16UC1
__global__ void k_convert_maps(const float* __restrict__ mx, const float* __restrict__ my, size_t n, short2* __restrict__ map1,
                               uint16_t* __restrict__ map2)
{
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int ix = __float2int_rn(__fmul_rn(mx[k], 32.f));
    const int iy = __float2int_rn(__fmul_rn(my[k], 32.f));
    map1[k] = make_short2((short)min(max(ix >> 5, -32768), 32767), (short)min(max(iy >> 5, -32768), 32767));
    map2[k] = (uint16_t)((iy & 31) * 32 + (ix & 31));
}

void remap_bilinear(const Launcher& L, const uint8_t* src, size_t sstep, int sH, int sW, const int16_t* map1,
                    const uint16_t* map2, int H, int W, uint8_t* dst, size_t dstep)
{
    dim3 b(RM_TW, RM_TH), g((W + RM_TW - 1) / RM_TW, (H + RM_TH - 1) / RM_TH);
    k_remap<<<g, b, 0, L.stream>>>(src, sstep, sH, sW, (const short2*)map1, map2, H, W, dst, dstep);
    L.count(1);
}

void convert_maps(const Launcher& L, const float* mx, const float* my, int H, int W, int16_t* map1, uint16_t* map2)
{
    const size_t n = (size_t)H * W;
    k_convert_maps<<<(unsigned)((n + 255) / 256), 256, 0, L.stream>>>(mx, my, n, (short2*)map1, map2);
    L.count(1);
}

// ---- rectify-map generation ------------------------------------------------------
// cv::initUndistortRectifyMap(K, D, R, P, size, CV_16SC2, map1, map2) as called by
// EpipolarRectifyMap::compute (source/stereo_utils.cpp:157-169): for every destination pixel (j, i)
//   [x y w]^T = (P_3x3 R)^-1 [j i 1]^T, x /= w, y /= w, r2 = x^2 + y^2,
//   kr = (1 + ((k3 r2 + k2) r2 + k1) r2) / (1 + ((k6 r2 + k5) r2 + k4) r2),
//   xd = x kr + p1 2xy + p2 (r2 + 2x^2) + s1 r2 + s2 r2^2,  yd likewise, tilt, u = fx xd + cx, v = fy yd + cy,
// then the fixed-point split of cv::convertMaps: iu = round(32 u), map1 = (iu >> 5, iv >> 5),
// map2 = (iv & 31) * 32 + (iu & 31).  All in fp64 like OpenCV; OpenCV's scalar loop accumulates
// x += ir[0] per column and its AVX2 path uses FMAs, so the last bits of u differ between OpenCV's own
// code paths -- only a value within ~1e-11 of a rounding tie of 32 u could differ after quantisation.
struct UndistortParams {
    double ir[9];      // (P_3x3 * R)^-1
    double k[14];      // k1 k2 p1 p2 k3 k4 k5 k6 s1 s2 s3 s4 tauX tauY
    double tilt[9];    // tilt projection matrix (identity when tauX = tauY = 0)
    double fx, fy, u0, v0;
};

__global__ void k_init_undistort_rectify_map(UndistortParams q, int H, int W, short2* __restrict__ map1, uint16_t* __restrict__ map2)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= W) return;
    const double _x = (double)i * q.ir[1] + q.ir[2] + (double)j * q.ir[0];
    const double _y = (double)i * q.ir[4] + q.ir[5] + (double)j * q.ir[3];
    const double _w = (double)i * q.ir[7] + q.ir[8] + (double)j * q.ir[6];
    const double w = 1. / _w, x = _x * w, y = _y * w;
    const double x2 = x * x, y2 = y * y;
    const double r2 = x2 + y2, _2xy = 2 * x * y;
    const double k1 = q.k[0], k2 = q.k[1], p1 = q.k[2], p2 = q.k[3], k3 = q.k[4], k4 = q.k[5], k5 = q.k[6], k6 = q.k[7];
    const double s1 = q.k[8], s2 = q.k[9], s3 = q.k[10], s4 = q.k[11];
    const double kr = (1 + ((k3 * r2 + k2) * r2 + k1) * r2) / (1 + ((k6 * r2 + k5) * r2 + k4) * r2);
    const double xd = (x * kr + p1 * _2xy + p2 * (r2 + 2 * x2) + s1 * r2 + s2 * r2 * r2);
    const double yd = (y * kr + p1 * (r2 + 2 * y2) + p2 * _2xy + s3 * r2 + s4 * r2 * r2);
    const double tx = q.tilt[0] * xd + q.tilt[1] * yd + q.tilt[2];
    const double ty = q.tilt[3] * xd + q.tilt[4] * yd + q.tilt[5];
    const double tz = q.tilt[6] * xd + q.tilt[7] * yd + q.tilt[8];
    const double inv = tz ? 1. / tz : 1;
    const double u = q.fx * inv * tx + q.u0;
    const double v = q.fy * inv * ty + q.v0;
    // saturate_cast<int>(double) = cvRound = round half to even, saturating
    auto sat_round = [](double t) {
        if (!(t > -2147483648.0)) return t != t ? (int)0x80000000 : (int)0x80000000;
        if (t >= 2147483647.0) return 2147483647;
        return __double2int_rn(t);
    };
    const int iu = sat_round(u * 32.0), iv = sat_round(v * 32.0);
    const size_t o = (size_t)i * W + j;
    map1[o] = make_short2((short)(iu >> 5), (short)(iv >> 5));
    map2[o] = (uint16_t)((iv & 31) * 32 + (iu & 31));
}

void init_undistort_rectify_map(const Launcher& L, const double* ir, const double* k14, const double* tilt, double fx, double fy,
                                double u0, double v0, int H, int W, int16_t* map1, uint16_t* map2)
{
    UndistortParams q;
    for (int a = 0; a < 9; ++a) { q.ir[a] = ir[a]; q.tilt[a] = tilt[a]; }
    for (int a = 0; a < 14; ++a) q.k[a] = k14[a];
    q.fx = fx; q.fy = fy; q.u0 = u0; q.v0 = v0;
    dim3 g((W + 127) / 128, H);
    k_init_undistort_rectify_map<<<g, 128, 0, L.stream>>>(q, H, W, reinterpret_cast<short2*>(map1), map2);
    L.count(1);
}

}  // namespace tsm
