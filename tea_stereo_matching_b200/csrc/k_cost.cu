// k_cost.cu -- AD + ternary-census initial cost volume (costInitialize,
// reference source/ADCensus.cpp:500-581).
//
//   C_k[d](y,x) = (2.f - expf(-ad/10.f)) - expf(-census/30.f)        ADCensus.cpp:518
//   view 0: (xL,xR) = (x, x-d);  view 1: (xL,xR) = (x+d, x)          ADCensus.cpp:556-561
//   2.f when xL+-4 / xR+-4 / y+-3 leaves the image                   ADCensus.cpp:562-566
//
// ad = (|dB|+|dG|+|dR|)/3.f has 766 distinct inputs, census 187: both exponentials
// come from tables built on the HOST with the host's expf (tsm_capi.cu), so the
// result is bit-identical to the reference's libm on whatever box it runs.
// census = sum_c popc((ltL_c & gtR_c) | (gtL_c & ltR_c))  (sign-product < 0 test, :469).
//
// Mapping: one warp per pixel, lanes over d (d = lane + 32k): the varying-view
// signature at x -/+ d is a (reversed) coalesced read, the fixed-view one a
// broadcast; the cost vector of a pixel is written as contiguous 128-byte rows.
#include "tsm_common.cuh"

namespace tsm {

constexpr int COST_PIX_PER_BLOCK = 64;
constexpr int COST_WARPS = 8;
constexpr int TAB_AD_N = 766, TAB_C_N = 192;

__global__ void __launch_bounds__(COST_WARPS * 32)
k_cost_init(Dims dm, ViewPtrs vl, ViewPtrs vr, const float* __restrict__ g_tab_ad, const float* __restrict__ g_tab_c)
{
    __shared__ float tab_ad[TAB_AD_N];
    __shared__ float tab_c[TAB_C_N];
    for (int i = threadIdx.x; i < TAB_AD_N; i += blockDim.x) tab_ad[i] = g_tab_ad[i];
    for (int i = threadIdx.x; i < TAB_C_N; i += blockDim.x) tab_c[i] = g_tab_c[i];
    __syncthreads();

    const int view = blockIdx.z;
    const int y = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int H = dm.H, W = dm.W, Dn = dm.Dn;
    const size_t npx = (size_t)H * W;
    const size_t row = (size_t)y * W;
    // fixed view = the view whose pixel stays at x; moving view is sampled at x -/+ d.
    const ViewPtrs& vf = view == 0 ? vl : vr;
    const ViewPtrs& vm = view == 0 ? vr : vl;
    const int sgn = view == 0 ? -1 : 1;
    const Vol vol = vf.vol;
    const bool yout = (y - kCensusH / 2 < 0) || (y + kCensusH / 2 >= H);
    const int hw = kCensusW / 2;

    const int xbeg = blockIdx.x * COST_PIX_PER_BLOCK;
    for (int x = xbeg + warp; x < min(xbeg + COST_PIX_PER_BLOCK, W); x += COST_WARPS) {
        float* out_main = vol.main + (row + x) * dm.Dm;
        float* out_tail = vol.tail + (row + x) * dm.Rp - dm.Dm;
        const bool fout = yout || (x - hw < 0) || (x + hw >= W);
        uint64_t fl[3], fg[3];
        uint32_t fpix = 0;
        if (!fout) {
            fpix = vf.img4[row + x];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                fl[c] = vf.census[(size_t)c * npx + row + x];
                fg[c] = vf.census[(size_t)(3 + c) * npx + row + x];
            }
        }
        for (int d = lane; d < Dn; d += 32) {
            const int xm = x + sgn * d;
            float cost = 2.f;
            if (!fout && xm - hw >= 0 && xm + hw < W) {
                const uint32_t mpix = vm.img4[row + xm];
                const int ad3 = __vsadu4(fpix, mpix);
                int cen = 0;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const uint64_t ml = vm.census[(size_t)c * npx + row + xm];
                    const uint64_t mg = vm.census[(size_t)(3 + c) * npx + row + xm];
                    cen += __popcll((fl[c] & mg) | (fg[c] & ml));
                }
                cost = __fsub_rn(__fsub_rn(2.f, tab_ad[ad3]), tab_c[cen]);
            }
            (d < dm.Dm ? out_main : out_tail)[d] = cost;
        }
    }
}

void cost_init(const Launcher& L, const Dims& d, const ViewPtrs& left, const ViewPtrs& right, const float* d_tab_ad,
               const float* d_tab_census)
{
    dim3 grid((d.W + COST_PIX_PER_BLOCK - 1) / COST_PIX_PER_BLOCK, d.H, 2);
    k_cost_init<<<grid, COST_WARPS * 32, 0, L.stream>>>(d, left, right, d_tab_ad, d_tab_census);
    L.count(1);
}

}  // namespace tsm
