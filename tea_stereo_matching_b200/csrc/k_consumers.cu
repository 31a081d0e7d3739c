// k_consumers.cu -- consumers of the disparity map: the free functions of the
// reference's source/stereo.cpp that run right after ADCensus::compute.
//   JETColorMap            stereo.cpp:75-93     (table, built on the host: jet_colormap)
//   applyColorMap          stereo.cpp:95-137    (auto range / explicit range)
//   reprojectToDepth       stereo.cpp:139-151   depth = (f*B) / d
//   reprojectTo3D (f,B)    stereo.cpp:153-172   Z = fB/d, X = (u-cx)*(Z/f), Y = (v-cy)*(Z/f)
//   reprojectTo3D (Q)      stereo.cpp:174-202   [X Y Z W]^T = Q32 * [u v d 1]^T (cv::gemm as OpenCV 4.13 does it:
//                                               fp32 products and sums, k ascending), then X/W, Y/W, Z/W
// Every expression keeps the reference's operation order in fp32 (explicit _rn intrinsics), so the maps are
// bit-identical to the CPU functions; pixels with d < 0 or d == +-inf stay zero as in the reference.
// One thread per pixel, all kernels are plain streaming (4 B in, 4 / 12 / 3 B out per pixel).
#include "tsm_common.cuh"
#include <math_constants.h>

namespace tsm {

__device__ __forceinline__ bool disp_skipped(float d) { return d < 0.f || isinf(d); }

__global__ void k_depth(const float* __restrict__ disp, float* __restrict__ depth, size_t n, float fb)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float d = disp[i];
    depth[i] = disp_skipped(d) ? 0.f : __fdiv_rn(fb, d);
}

__global__ void k_xyz_fb(const float* __restrict__ disp, float* __restrict__ xyz, int H, int W, float fb, float f, float cx, float cy)
{
    const int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y;
    if (u >= W) return;
    const size_t i = (size_t)v * W + u;
    const float d = disp[i];
    float X = 0.f, Y = 0.f, Z = 0.f;
    if (!disp_skipped(d)) {
        Z = __fdiv_rn(fb, d);
        const float zf = __fdiv_rn(Z, f);
        X = __fmul_rn(__fsub_rn((float)u, cx), zf);
        Y = __fmul_rn(__fsub_rn((float)v, cy), zf);
    }
    xyz[3 * i + 0] = X;
    xyz[3 * i + 1] = Y;
    xyz[3 * i + 2] = Z;
}

struct QMat { float q[16]; };

__global__ void k_xyz_q(const float* __restrict__ disp, float* __restrict__ xyz, int H, int W, QMat Q)
{
    const int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y;
    if (u >= W) return;
    const size_t i = (size_t)v * W + u;
    const float p[4] = {(float)u, (float)v, disp[i], 1.f};
    float r[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        float s = 0.f;  // cv::gemm (OpenCV 4.13, 4x4 by 4xN): fp32 products and sums, k ascending, no FMA
#pragma unroll
        for (int k = 0; k < 4; ++k) s = __fadd_rn(s, __fmul_rn(Q.q[4 * a + k], p[k]));
        r[a] = s;
    }
    xyz[3 * i + 0] = __fdiv_rn(r[0], r[3]);
    xyz[3 * i + 1] = __fdiv_rn(r[1], r[3]);
    xyz[3 * i + 2] = __fdiv_rn(r[2], r[3]);
}

// min / max over the values the reference looks at (not < 0, not inf; NaN never wins a std::min / std::max).
// Keys are the int bit patterns of non-negative floats (order-preserving); range[0] starts at +inf,
// range[1] at -1 = "no value" (decoded as -inf).
__global__ void __launch_bounds__(256) k_minmax(const float* __restrict__ disp, size_t n, int* __restrict__ range)
{
    int lo = 0x7f800000, hi = -1;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float d = disp[i];
        if (d < 0.f || isinf(d) || d != d) continue;
        const int k = __float_as_int(d) & 0x7fffffff;  // -0.0 behaves like +0.0 in every later expression
        lo = min(lo, k);
        hi = max(hi, k);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&range[0], lo);
        atomicMax(&range[1], hi);
    }
}

// static_cast<unsigned char>(float) as x86-64 compiles it: cvttss2si (INT_MIN for NaN / out of range), low byte
__device__ __forceinline__ unsigned char to_uchar_x86(float t)
{
    int i;
    if (t != t || t >= 2147483648.f || t < -2147483648.f) i = INT_MIN;
    else i = (int)t;  // truncation
    return (unsigned char)(i & 0xff);
}

__global__ void k_colormap(const float* __restrict__ disp, uint8_t* __restrict__ dst, size_t n, int auto_range,
                           const int* __restrict__ range, float minv, float maxv, const uint8_t* __restrict__ table)
{
    __shared__ uint8_t tab[768];
    for (int i = threadIdx.x; i < 768; i += blockDim.x) tab[i] = table[i];
    __syncthreads();
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float mn = minv, mx = maxv;
    if (auto_range) {
        mn = __int_as_float(range[0]);
        mx = range[1] < 0 ? -CUDART_INF_F : __int_as_float(range[1]);
    }
    const float d = disp[i];
    const bool black = auto_range ? (d < 0.f) : (d < mn || d > mx);
    uint8_t b = 0, g = 0, r = 0;
    if (!black) {
        const float t = __fmul_rn(__fdiv_rn(__fsub_rn(d, mn), __fsub_rn(mx, mn)), 255.f);
        const int idx = to_uchar_x86(t);
        b = tab[3 * idx]; g = tab[3 * idx + 1]; r = tab[3 * idx + 2];
    }
    dst[3 * i] = b; dst[3 * i + 1] = g; dst[3 * i + 2] = r;
}

void jet_colormap(uint8_t* t)
{
    auto set = [&](int i, int b, int g, int r) { t[3 * i] = (uint8_t)b; t[3 * i + 1] = (uint8_t)g; t[3 * i + 2] = (uint8_t)r; };
    for (int i = 0; i < 32; ++i) set(i, 128 + 4 * i, 0, 0);
    set(32, 255, 0, 0);
    for (int i = 0; i < 63; ++i) set(33 + i, 255, 4 + 4 * i, 0);
    set(96, 254, 255, 2);
    for (int i = 0; i < 62; ++i) set(97 + i, 250 - 4 * i, 255, 6 + 4 * i);
    set(159, 1, 255, 254);
    for (int i = 0; i < 64; ++i) set(160 + i, 0, 252 - 4 * i, 255);
    for (int i = 0; i < 32; ++i) set(224 + i, 0, 0, 252 - 4 * i);
}

void reproject_depth(const Launcher& L, const float* disp, float* depth, size_t n, float focal, float baseline)
{
    const float fb = focal * baseline;  // stereo.cpp:144
    k_depth<<<(unsigned)((n + 255) / 256), 256, 0, L.stream>>>(disp, depth, n, fb);
    L.count(1);
}

void reproject_xyz_fb(const Launcher& L, const float* disp, float* xyz, int H, int W, float focal, float baseline, float cx, float cy)
{
    dim3 g((W + 255) / 256, H);
    k_xyz_fb<<<g, 256, 0, L.stream>>>(disp, xyz, H, W, focal * baseline, focal, cx, cy);
    L.count(1);
}

void reproject_xyz_q(const Launcher& L, const float* disp, float* xyz, int H, int W, const double* Q)
{
    QMat q;
    for (int i = 0; i < 16; ++i) q.q[i] = (float)Q[i];  // Q.convertTo(CV_32F), stereo.cpp:190-191
    dim3 g((W + 255) / 256, H);
    k_xyz_q<<<g, 256, 0, L.stream>>>(disp, xyz, H, W, q);
    L.count(1);
}

__global__ void k_range_init(int* range)
{
    range[0] = 0x7f800000;  // +inf as the running minimum
    range[1] = -1;          // below every non-negative float pattern as the running maximum
}

void apply_colormap(const Launcher& L, const float* disp, uint8_t* dst, size_t n, bool auto_range, float minv, float maxv,
                    const uint8_t* d_table, int* d_range)
{
    if (auto_range) {
        k_range_init<<<1, 1, 0, L.stream>>>(d_range);
        L.count(1);
        const unsigned blocks = (unsigned)std::min<size_t>((n + 255) / 256, 148 * 8);
        k_minmax<<<blocks, 256, 0, L.stream>>>(disp, n, d_range);
        L.count(1);
    }
    k_colormap<<<(unsigned)((n + 255) / 256), 256, 0, L.stream>>>(disp, dst, n, auto_range ? 1 : 0, d_range, minv, maxv, d_table);
    L.count(1);
}

}  // namespace tsm
