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
// Mapping: one CTA = (view, row y, 64-pixel tile).  The signatures + BGRx pixels of the
// MOVING view (the one sampled at x -/+ d) for the tile's whole disparity reach are staged
// once in shared memory.  A warp handles 4 consecutive pixels of the fixed view at a time;
// lane l of chunk k owns ONE moving column and evaluates it against the 4 fixed pixels
// (disparities e, e+1, e+2, e+3 -- resp. e, e-1, ... for the right view), so each staged
// signature is read once per 4 cells and every pixel's cost vector is still written as
// lane-contiguous rows.  The six 32-bit match words of a cell are summed with a carry-save
// adder tree (3 POPC instead of 6: POPC is a quarter-rate instruction).
#include "tsm_common.cuh"

namespace tsm {

constexpr int COST_TX = 64;     // pixels per CTA
constexpr int COST_WARPS = 8;
constexpr int COST_J = 4;       // fixed pixels per warp iteration
constexpr int TAB_AD_N = 766, TAB_C_N = 192;
constexpr int TAB_C_PAD = 194;  // (766 + 194) * 4 bytes = 3840: keeps what follows 16-byte aligned
constexpr uint32_t kInvalidPix = 0xffffffffu;

struct Sig {
    uint32_t w[12];  // lt_B lo,hi, lt_G lo,hi, lt_R lo,hi, gt_B lo,hi, gt_G lo,hi, gt_R lo,hi
    uint32_t pix;    // BGRx, or kInvalidPix when the census window leaves the image
};

__device__ __forceinline__ int census_count(const Sig& f, const Sig& m)
{
    // match word i = (lt_f & gt_m) | (gt_f & lt_m) for the six 32-bit halves
    uint32_t a[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) a[i] = (f.w[i] & m.w[6 + i]) | (f.w[6 + i] & m.w[i]);
    // carry-save adders: popc(a0..a5) = popc(ones) + 2 popc(twos) + 4 popc(fours)
    const uint32_t s1 = a[0] ^ a[1] ^ a[2], c1 = (a[0] & a[1]) | (a[2] & (a[0] ^ a[1]));
    const uint32_t s2 = a[3] ^ a[4] ^ a[5], c2 = (a[3] & a[4]) | (a[5] & (a[3] ^ a[4]));
    const uint32_t ones = s1 ^ s2, t = s1 & s2;
    const uint32_t twos = c1 ^ c2 ^ t, fours = (c1 & c2) | (t & (c1 ^ c2));
    return __popc(ones) + 2 * __popc(twos) + 4 * __popc(fours);
}

__global__ void __launch_bounds__(COST_WARPS * 32)
k_cost_init(Dims dm, ViewPtrs vl, ViewPtrs vr, const float* __restrict__ g_tab_ad, const float* __restrict__ g_tab_c)
{
    extern __shared__ __align__(16) uint32_t smem[];
    const int view = blockIdx.z, y = blockIdx.y, x0 = blockIdx.x * COST_TX;
    const int H = dm.H, W = dm.W, Dn = dm.Dn;
    const int ncol = (COST_TX + Dn - 1 + 3) & ~3;  // moving columns the tile can reach (padded to keep 16-byte alignment)
    float* tab_ad = reinterpret_cast<float*>(smem);
    float* tab_c = tab_ad + TAB_AD_N;
    uint32_t* mw = reinterpret_cast<uint32_t*>(tab_c + TAB_C_PAD);  // [ncol][13]: 12 signature words + pixel (stride 13: conflict-free)
    for (int i = threadIdx.x; i < TAB_AD_N; i += blockDim.x) tab_ad[i] = g_tab_ad[i];
    for (int i = threadIdx.x; i < TAB_C_N; i += blockDim.x) tab_c[i] = g_tab_c[i];

    const ViewPtrs& vf = view == 0 ? vl : vr;  // fixed view: its pixel stays at x
    const ViewPtrs& vm = view == 0 ? vr : vl;  // moving view: sampled at x - d (view 0) / x + d (view 1)
    const size_t npx = (size_t)H * W, row = (size_t)y * W;
    const int hw = kCensusW / 2;
    const bool yout = (y - kCensusH / 2 < 0) || (y + kCensusH / 2 >= H);
    // first staged moving column: view 0 reaches down to x0 - (Dn-1); view 1 starts at x0
    const int cbase = view == 0 ? x0 - (Dn - 1) : x0;
    for (int i = threadIdx.x; i < ncol; i += blockDim.x) {
        const int c = cbase + i;
        const bool ok = !yout && c - hw >= 0 && c + hw < W;
        mw[13 * i + 12] = ok ? vm.img4[row + c] : kInvalidPix;
#pragma unroll
        for (int p = 0; p < 6; ++p) {
            const uint64_t s = ok ? vm.census[(size_t)p * npx + row + c] : 0ull;
            mw[13 * i + 2 * p] = (uint32_t)s;
            mw[13 * i + 2 * p + 1] = (uint32_t)(s >> 32);
        }
    }
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const Vol vol = vf.vol;
    const int nchunk = (Dn + COST_J - 1 + 31) / 32;
    const int DnP = (Dn + 3) & ~3;
    float* wtile = reinterpret_cast<float*>(mw + 13 * ncol);  // [COST_WARPS][COST_J][DnP], 16-byte aligned
    for (int g = warp; g < COST_TX / COST_J; g += COST_WARPS) {
        const int x = x0 + g * COST_J;
        if (x >= W) break;
        // the four fixed pixels x .. x+3 (uniform across the warp)
        Sig f[COST_J];
#pragma unroll
        for (int j = 0; j < COST_J; ++j) {
            const int xf = x + j;
            const bool ok = !yout && xf < W && xf - hw >= 0 && xf + hw < W;
            f[j].pix = ok ? vf.img4[row + xf] : kInvalidPix;
#pragma unroll
            for (int p = 0; p < 6; ++p) {
                const uint64_t s = ok ? vf.census[(size_t)p * npx + row + xf] : 0ull;
                f[j].w[2 * p] = (uint32_t)s;
                f[j].w[2 * p + 1] = (uint32_t)(s >> 32);
            }
        }
        bool fok[COST_J];
#pragma unroll
        for (int j = 0; j < COST_J; ++j) {
            fok[j] = f[j].pix != kInvalidPix;
            // keep the fixed signatures in registers: without this the compiler re-loads all 28 words
            // from global memory in every chunk iteration instead of keeping 52 registers live
#pragma unroll
            for (int i = 0; i < 12; ++i) asm volatile("" : "+r"(f[j].w[i]));
            asm volatile("" : "+r"(f[j].pix));
        }
        // Costs go to a per-warp shared tile [J][DnP] first: in this mapping a lane's disparity for
        // pixel j is e +- j, so direct global stores would all be sector-misaligned partial writes.
        float* tile = wtile + warp * (COST_J * DnP);
        float* tj[COST_J];  // tj[j] + e addresses disparity d_j = e +- j of pixel j
#pragma unroll
        for (int j = 0; j < COST_J; ++j) tj[j] = tile + j * DnP + (view == 0 ? j : -j);
        for (int k = 0; k < nchunk; ++k) {
            // e = disparity of this lane's moving column against fixed pixel j = 0
            //   view 0: column c = x - e, d_j = e + j, e in [-(J-1), Dn-1]
            //   view 1: column c = x + e, d_j = e - j, e in [0, Dn-1+J-1]
            const int e = 32 * k + lane - (view == 0 ? COST_J - 1 : 0);
            const int ci = (view == 0 ? x - e : x + e) - cbase;
            const bool cin = ci >= 0 && ci < ncol;
            const int cs = cin ? ci : 0;
            Sig m;
            const uint32_t* mc = mw + 13 * cs;
            m.pix = mc[12];
#pragma unroll
            for (int i = 0; i < 12; ++i) m.w[i] = mc[i];
            const bool mok = cin && m.pix != kInvalidPix;
#pragma unroll
            for (int j = 0; j < COST_J; ++j) {
                const int d = view == 0 ? e + j : e - j;
                const int ad3 = min((int)__vsadu4(f[j].pix, m.pix), TAB_AD_N - 1);
                const int cen = census_count(f[j], m);
                float cost = __fsub_rn(__fsub_rn(2.f, tab_ad[ad3]), tab_c[cen]);
                cost = (fok[j] && mok) ? cost : 2.f;
                if ((unsigned)d < (unsigned)Dn) tj[j][e] = cost;
            }
        }
        __syncwarp();
        // flush: 16-byte aligned vector stores of the main part, scalar stores of the tail part
        const int nq = dm.Dm / 4;  // float4 per pixel in the main part
        const int r = dm.tail();
#pragma unroll
        for (int j = 0; j < COST_J; ++j) {
            if (x + j < W) {
                const float4* src = reinterpret_cast<const float4*>(tile + j * DnP);
                float4* dst = reinterpret_cast<float4*>(vol.main + (row + x + j) * dm.Dm);
                for (int q = lane; q < nq; q += 32) dst[q] = src[q];
                if (lane < r) vol.tail[(row + x + j) * dm.Rp + lane] = tile[j * DnP + dm.Dm + lane];
            }
        }
        __syncwarp();
    }
}

void cost_init(const Launcher& L, const Dims& d, const ViewPtrs& left, const ViewPtrs& right, const float* d_tab_ad,
               const float* d_tab_census)
{
    const size_t ncol = (size_t)((COST_TX + d.Dn - 1 + 3) & ~3), dnp = (size_t)((d.Dn + 3) & ~3);
    const size_t smem = (size_t)(TAB_AD_N + TAB_C_PAD) * 4 + 13 * ncol * 4 + (size_t)COST_WARPS * COST_J * dnp * 4;
    static size_t smem_set = 0;
    if (smem > smem_set) {
        cudaFuncSetAttribute(k_cost_init, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        smem_set = smem;
    }
    dim3 grid((d.W + COST_TX - 1) / COST_TX, d.H, 2);
    k_cost_init<<<grid, COST_WARPS * 32, smem, L.stream>>>(d, left, right, d_tab_ad, d_tab_census);
    L.count(1);
}

}  // namespace tsm
