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

}  // namespace tsm
