// Does a "blocked" store (lane l writes 6 consecutive floats of a 768-byte vector: 3 x STG.64 at a 24-byte lane stride)
// run at the HBM rate, or does it need staging?  Copy kernel, one 768-byte vector per warp step.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256) cp(const float* __restrict__ in, float* __restrict__ out, size_t nvec)
{
    const int lane = threadIdx.x & 31;
    size_t w = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const size_t nw = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (; w < nvec; w += nw) {
        const float* s = in + w * 192;
        float* d = out + w * 192;
        if (MODE == 0) {  // striped in, striped out
            float v[6];
#pragma unroll
            for (int k = 0; k < 6; ++k) v[k] = s[lane + 32 * k];
#pragma unroll
            for (int k = 0; k < 6; ++k) d[lane + 32 * k] = v[k] + 1.f;
        } else if (MODE == 1) {  // blocked in (3 x LDG.64), blocked out (3 x STG.64)
            float2 v[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) v[k] = reinterpret_cast<const float2*>(s + 6 * lane)[k];
#pragma unroll
            for (int k = 0; k < 3; ++k) reinterpret_cast<float2*>(d + 6 * lane)[k] = make_float2(v[k].x + 1.f, v[k].y + 1.f);
        } else {  // striped in, blocked out
            float v[6];
#pragma unroll
            for (int k = 0; k < 6; ++k) v[k] = s[lane + 32 * k];
#pragma unroll
            for (int k = 0; k < 3; ++k) reinterpret_cast<float2*>(d + 6 * lane)[k] = make_float2(v[2 * k] + 1.f, v[2 * k + 1] + 1.f);
        }
    }
}
int main()
{
    const size_t bytes = (size_t)3072 << 20, nvec = bytes / 768;
    float *a, *b;
    cudaMalloc(&a, bytes); cudaMalloc(&b, bytes); cudaMemset(a, 0, bytes); cudaMemset(b, 0, bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float ms;
    for (int g : {148 * 8, 148 * 32}) {
        for (int m = 0; m < 3; ++m) {
            for (int r = 0; r < 2; ++r) {
                cudaEventRecord(e0);
                if (m == 0) cp<0><<<g, 256>>>(a, b, nvec); else if (m == 1) cp<1><<<g, 256>>>(a, b, nvec); else cp<2><<<g, 256>>>(a, b, nvec);
                cudaEventRecord(e1); cudaDeviceSynchronize(); cudaEventElapsedTime(&ms, e0, e1);
            }
            printf("grid=%5d mode=%d (%s): %.3f ms  %.0f GB/s (read+write)\n", g, m,
                   m == 0 ? "striped/striped" : m == 1 ? "blocked/blocked" : "striped/blocked", ms, 2.0 * bytes / ms / 1e6);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
