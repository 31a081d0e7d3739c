#include <cstdio>
#include <cuda_runtime.h>
__global__ void wr4(float4* p, size_t n) { size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; const size_t st = (size_t)gridDim.x * blockDim.x; for (; i < n; i += st) p[i] = make_float4(1.f, 2.f, 3.f, (float)i); }
__global__ void wr1(float* p, size_t n) { size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; const size_t st = (size_t)gridDim.x * blockDim.x; for (; i < n; i += st) p[i] = (float)i; }
__global__ void rd4(const float4* p, size_t n, float* out) { size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; const size_t st = (size_t)gridDim.x * blockDim.x; float a = 0; for (; i < n; i += st) { float4 v = p[i]; a += v.x + v.y + v.z + v.w; } if (a == 123.f) *out = a; }
int main() {
    const size_t bytes = (size_t)6400 << 20; float* p; cudaMalloc(&p, bytes); cudaMemset(p, 0, bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float ms;
    for (int g : {148 * 4, 148 * 16, 148 * 64}) {
        wr4<<<g, 256>>>((float4*)p, bytes / 16); cudaDeviceSynchronize();
        cudaEventRecord(e0); wr4<<<g, 256>>>((float4*)p, bytes / 16); cudaEventRecord(e1); cudaDeviceSynchronize();
        cudaEventElapsedTime(&ms, e0, e1); printf("write float4 grid=%5d: %.3f ms %.0f GB/s\n", g, ms, bytes / ms / 1e6);
        cudaEventRecord(e0); wr1<<<g, 256>>>(p, bytes / 4); cudaEventRecord(e1); cudaDeviceSynchronize();
        cudaEventElapsedTime(&ms, e0, e1); printf("write float  grid=%5d: %.3f ms %.0f GB/s\n", g, ms, bytes / ms / 1e6);
        cudaEventRecord(e0); rd4<<<g, 256>>>((const float4*)p, bytes / 16, p); cudaEventRecord(e1); cudaDeviceSynchronize();
        cudaEventElapsedTime(&ms, e0, e1); printf("read  float4 grid=%5d: %.3f ms %.0f GB/s\n", g, ms, bytes / ms / 1e6);
    }
    cudaEventRecord(e0); cudaMemsetAsync(p, 1, bytes); cudaEventRecord(e1); cudaDeviceSynchronize();
    cudaEventElapsedTime(&ms, e0, e1); printf("cudaMemset: %.3f ms %.0f GB/s\n", ms, bytes / ms / 1e6);
    return 0;
}
