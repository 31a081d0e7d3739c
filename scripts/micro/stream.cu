// Microbenchmark: how much HBM bandwidth can the aggregation access pattern reach with few warps per SM?
// Each warp walks a line of pixel vectors (stride Dp floats), reading VEC floats per lane per step and
// writing the same cell LAG steps later.  Occupancy is limited with dummy dynamic shared memory.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

template <int VEC> struct V;
template <> struct V<2> { typedef float2 T; };
template <> struct V<4> { typedef float4 T; };
__device__ inline float2 bump(float2 v) { v.x += 1.f; v.y += 1.f; return v; }
__device__ inline float4 bump(float4 v) { v.x += 1.f; v.y += 1.f; v.z += 1.f; v.w += 1.f; return v; }

template <int VEC, int PF, bool VERT>
__global__ void walk(float* vol, int H, int W, int Dp, int chunks_per_line, int lag, const unsigned* small)
{
    extern __shared__ unsigned char dummy[];
    typedef typename V<VEC>::T T;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int nlines = VERT ? W : H, len = VERT ? H : W;
    if (warp >= nlines * chunks_per_line) return;
    const int line = warp / chunks_per_line, chunk = warp % chunks_per_line;
    const int d = chunk * 32 * VEC + lane * VEC;
    if (d + VEC > Dp) return;
    const size_t stride = VERT ? (size_t)W * Dp : (size_t)Dp;
    float* base = vol + (VERT ? (size_t)line * Dp : (size_t)line * W * Dp) + d;
    const float* in = base;
    float* out = base;
    const unsigned* sm = small ? small + (VERT ? (size_t)line : (size_t)line * W) : nullptr;
    const size_t sstride = VERT ? W : 1;
    unsigned acc = 0;
    // extra lag: read `lag` positions ahead before the first write
    for (int t = 0; t < lag; ++t) { acc += __float_as_uint(*(in + t * stride)); }
    in += (size_t)lag * stride;
    T buf[2][PF];
#pragma unroll
    for (int u = 0; u < PF; ++u) buf[0][u] = *reinterpret_cast<const T*>(in + u * stride);
    in += PF * stride;
    int t = 0;
    for (; t + 2 * PF + lag <= len; t += 2 * PF) {
        if (sm) {
#pragma unroll
            for (int u = 0; u < 2 * PF; ++u) acc += sm[(size_t)(t + u) * sstride];
        }
#pragma unroll
        for (int u = 0; u < PF; ++u) buf[1][u] = *reinterpret_cast<const T*>(in + u * stride);
        in += PF * stride;
#pragma unroll
        for (int u = 0; u < PF; ++u) { *reinterpret_cast<T*>(out) = bump(buf[0][u]); out += stride; }
#pragma unroll
        for (int u = 0; u < PF; ++u) buf[0][u] = *reinterpret_cast<const T*>(in + u * stride);
        in += PF * stride;
#pragma unroll
        for (int u = 0; u < PF; ++u) { *reinterpret_cast<T*>(out) = bump(buf[1][u]); out += stride; }
    }
    if ((dummy[0] == 123 && lane == 77) || acc == 0x12345u) out[0] = 0;
}

template <int VEC, int PF, bool VERT>
void run(float* vol, int H, int W, int Dp, int warps_per_sm, const char* tag, int bt = 64, int lag = 0, const unsigned* small = nullptr)
{
    const int chunks = (Dp + 32 * VEC - 1) / (32 * VEC);
    const int nwarps = (VERT ? W : H) * chunks;
    const int blocks_per_sm = warps_per_sm / (bt / 32);
    size_t smem = (227 * 1024) / blocks_per_sm - 1024;
    smem &= ~(size_t)127;
    CK(cudaFuncSetAttribute(walk<VEC, PF, VERT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (nwarps * 32 + bt - 1) / bt;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    walk<VEC, PF, VERT><<<grid, bt, smem>>>(vol, H, W, Dp, chunks, lag, small);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    for (int i = 0; i < 3; ++i) walk<VEC, PF, VERT><<<grid, bt, smem>>>(vol, H, W, Dp, chunks, lag, small);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 3;
    const double bytes = 2.0 * H * W * (double)Dp * 4;
    printf("%-6s lag=%2d small=%d Dp=%d bt=%3d VEC=%d PF=%2d warps/SM=%2d %s: %.3f ms  %.0f GB/s\n", tag, lag, small != nullptr, Dp, bt, VEC, PF, warps_per_sm, VERT ? "V" : "H", ms, bytes / ms / 1e6);
}

int main()
{
    const int H = 1080, W = 1920, Dp = 224;
    float* vol;
    const size_t n = (size_t)H * W * Dp + (size_t)64 * W * Dp;
    CK(cudaMalloc(&vol, n * 4));
    CK(cudaMemset(vol, 0, n * 4));
    unsigned* small;
    CK(cudaMalloc(&small, (size_t)H * W * 4));
    CK(cudaMemset(small, 1, (size_t)H * W * 4));
    const int dp = 192;
    run<2, 8, false>(vol, H, W, dp, 6, "walk", 64, 33, nullptr);
    run<2, 8, false>(vol, H, W + 1, dp, 6, "walk", 64, 33, nullptr);    // line pitch + 1 pixel: partition camping?
    run<2, 8, false>(vol, H, W + 5, dp, 6, "walk", 64, 33, nullptr);
    run<2, 8, true>(vol, H, W + 1, dp, 6, "walk", 64, 33, nullptr);
    // copy-like reference in the same harness: "pixel vector" = 64 floats -> every warp streams contiguously
    run<2, 8, false>(vol, 1080 * 3, W, 64, 6, "copy", 64, 33, nullptr);
    run<2, 8, false>(vol, 1080 * 3, W, 64, 24, "copy", 64, 33, nullptr);
    run<4, 8, false>(vol, 1080 * 3 / 2, W, 128, 24, "copy", 64, 33, nullptr);
    run<2, 8, false>(vol, H, W, dp, 24, "walk", 64, 33, nullptr);
    run<2, 8, false>(vol, H, W, dp, 48, "walk", 64, 33, nullptr);
    return 0;
}
