// k_scanline.cu -- four cascaded in-place scanline passes (scanlineOptimize,
// reference source/ADCensus.cpp:795-1011): down, up, right, left, per view.
//
// For pixel p with predecessor q on the path (partialOptimization, :869-913):
//   m = min_k C(q,k);  if m == 0 the pixel is skipped                         (:871-881)
//   C(p,d) <- ((C(p,d) - m) + min(C(q,d), C(q,d-1)+P1, C(q,d+1)+P1, m+P2)) / 2  (:883-911)
// (P1,P2) by the number s of similar image pairs (computeP1P2, :915-981):
//   own view: colorDiff(I(p), I(q)) < 15; other view at columns x_p + sgn*d, x_q + sgn*d
//   (sgn = +1 for the left volume, -1 for the right one -- the reference's sign, :919-934),
//   counted only when both columns are inside the image;
//   s = 2 -> (1, 3); s = 1 -> (1/4, 3/4); s = 0 -> (1/10.f, 3/10.f).
// Only - + min and an exact halving are involved, so any evaluation order over d is
// bit-identical to the reference's sequential loop.
//
// Mapping: one warp owns a line (a column for the vertical passes, a row for the
// horizontal ones) and walks it forward then backward; the predecessor's updated cost
// vector stays in registers (lanes over d, d = lane + 32k), min over d is a shuffle
// butterfly, the d-1 / d+1 neighbours come from two rotate-shuffles per register.
// The inputs of the next pixels do not depend on the recurrence and are loaded a batch
// ahead.  The pass pair (down+up, right+left) is one launch; both views share a launch.
#include "tsm_common.cuh"
#include <math_constants.h>

namespace tsm {

constexpr int SCAN_WARPS = 4;
constexpr int SCAN_U = 4;  // steps per batch

struct ScanParams {
    float p1[3];
    float p2[3];
};

template <int K>
__device__ __forceinline__ void scan_step(float (&prev)[K], const float (&cur)[K], unsigned oth, int own,
                                          float* __restrict__ out, int Dn, int lane, const ScanParams& sp)
{
    float m = prev[0];
#pragma unroll
    for (int k = 1; k < K; ++k) m = fminf(m, prev[k]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (m == 0.f) {  // ADCensus.cpp:880 -- pixel left untouched
#pragma unroll
        for (int k = 0; k < K; ++k) prev[k] = cur[k];
        return;
    }
    float rl[K], rr[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        rl[k] = __shfl_sync(0xffffffffu, prev[k], (lane + 31) & 31);
        rr[k] = __shfl_sync(0xffffffffu, prev[k], (lane + 1) & 31);
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int d = lane + 32 * k;
        const float lo = (lane == 0) ? (k > 0 ? rl[k > 0 ? k - 1 : 0] : CUDART_INF_F) : rl[k];
        const float hi = (lane == 31) ? (k < K - 1 ? rr[k < K - 1 ? k + 1 : K - 1] : CUDART_INF_F) : rr[k];
        const int s = own + (int)((oth >> k) & 1u);
        const float p1 = s == 2 ? sp.p1[2] : (s == 1 ? sp.p1[1] : sp.p1[0]);
        const float p2 = s == 2 ? sp.p2[2] : (s == 1 ? sp.p2[1] : sp.p2[0]);
        float mo = __fadd_rn(m, p2);
        mo = fminf(mo, prev[k]);
        mo = fminf(mo, __fadd_rn(lo, p1));
        mo = fminf(mo, __fadd_rn(hi, p1));
        const float nv = __fmul_rn(__fadd_rn(__fsub_rn(cur[k], m), mo), 0.5f);
        if (d < Dn) {
            out[d] = nv;
            prev[k] = nv;
        } else {
            prev[k] = CUDART_INF_F;
        }
    }
}

// One direction of one line.  pos runs from `first` by `step` for `count` pixels; the
// predecessor of pos is pos - step.  Flag geometry (see file header):
//   VERT : flag row = row of max(pos, pred) (bit 0), column offset 0, valid columns [0, W-1]
//   HORZ : flag col = max(pos, pred) (bit 1): forward offset 0 valid [1, W-1]; backward offset 1 valid [0, W-2]
template <int K, bool VERT>
__device__ __forceinline__ void scan_dir(float (&prev)[K], float* __restrict__ vol, const uint8_t* __restrict__ fown,
                                         const uint8_t* __restrict__ foth, const Dims& dm, int line, int first, int step,
                                         int count, int sgn, int lane, const ScanParams& sp)
{
    const int W = dm.W, Dn = dm.Dn, Dp = dm.Dp;
    const int bit = VERT ? 0 : 1;
    const int coff = (!VERT && step < 0) ? 1 : 0;
    const int lo = (!VERT && step > 0) ? 1 : 0;
    const int hi = (!VERT && step < 0) ? W - 2 : W - 1;
    for (int i0 = 0; i0 < count; i0 += SCAN_U) {
        float cur[SCAN_U][K];
        unsigned oth[SCAN_U];
        int own[SCAN_U];
#pragma unroll
        for (int u = 0; u < SCAN_U; ++u) {
            const int i = i0 + u;
            if (i < count) {
                const int pos = first + i * step;
                const int y = VERT ? pos : line, x = VERT ? line : pos;
                const int fy = VERT ? (step > 0 ? pos : pos + 1) : line;
                const float* src = vol + ((size_t)y * W + x) * Dp;
                const uint8_t* frow_own = fown + (size_t)fy * W;
                const uint8_t* frow_oth = foth + (size_t)fy * W;
                own[u] = (frow_own[x + coff] >> bit) & 1;
                oth[u] = 0u;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const int d = lane + 32 * k;
                    cur[u][k] = d < Dn ? src[d] : CUDART_INF_F;
                    const int xo = x + sgn * d;
                    if (d < Dn && xo >= lo && xo <= hi) oth[u] |= (unsigned)((frow_oth[xo + coff] >> bit) & 1) << k;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < SCAN_U; ++u) {
            const int i = i0 + u;
            if (i < count) {
                const int pos = first + i * step;
                const int y = VERT ? pos : line, x = VERT ? line : pos;
                scan_step<K>(prev, cur[u], oth[u], own[u], vol + ((size_t)y * W + x) * Dp, Dn, lane, sp);
            }
        }
    }
}

template <int K, bool VERT>
__global__ void __launch_bounds__(SCAN_WARPS * 32)
k_scanline(Dims dm, ViewPtrs v0, ViewPtrs v1, ScanParams sp)
{
    const int view = blockIdx.y;
    const ViewPtrs& v = view ? v1 : v0;
    const ViewPtrs& o = view ? v0 : v1;
    const int lane = threadIdx.x & 31;
    const int line = blockIdx.x * SCAN_WARPS + (threadIdx.x >> 5);
    const int nlines = VERT ? dm.W : dm.H, len = VERT ? dm.H : dm.W;
    if (line >= nlines) return;
    const int sgn = view == 0 ? 1 : -1;
    float prev[K];
    {
        const float* src = v.vol + (VERT ? (size_t)line : (size_t)line * dm.W) * dm.Dp;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int d = lane + 32 * k;
            prev[k] = d < dm.Dn ? src[d] : CUDART_INF_F;
        }
    }
    // forward: pos = 1 .. len-1 (pred pos-1); backward: pos = len-2 .. 0 (pred pos+1).
    scan_dir<K, VERT>(prev, v.vol, v.flags, o.flags, dm, line, 1, 1, len - 1, sgn, lane, sp);
    scan_dir<K, VERT>(prev, v.vol, v.flags, o.flags, dm, line, len - 2, -1, len - 1, sgn, lane, sp);
}

template <int K>
static void launch_scan(const Launcher& L, const Dims& d, const ViewPtrs& left, const ViewPtrs& right, const ScanParams& sp)
{
    dim3 gv((d.W + SCAN_WARPS - 1) / SCAN_WARPS, 2), gh((d.H + SCAN_WARPS - 1) / SCAN_WARPS, 2);
    k_scanline<K, true><<<gv, SCAN_WARPS * 32, 0, L.stream>>>(d, left, right, sp);
    k_scanline<K, false><<<gh, SCAN_WARPS * 32, 0, L.stream>>>(d, left, right, sp);
    L.count(2);
}

void scanline(const Launcher& L, const Dims& d, const ViewPtrs& left, const ViewPtrs& right, float p1_lo, float p2_lo)
{
    ScanParams sp;
    sp.p1[0] = p1_lo; sp.p1[1] = 0.25f; sp.p1[2] = 1.f;
    sp.p2[0] = p2_lo; sp.p2[1] = 0.75f; sp.p2[2] = 3.f;
    const int K = (d.Dn + 31) / 32;
    switch (K) {
        case 1: launch_scan<1>(L, d, left, right, sp); break;
        case 2: launch_scan<2>(L, d, left, right, sp); break;
        case 3: launch_scan<3>(L, d, left, right, sp); break;
        case 4: launch_scan<4>(L, d, left, right, sp); break;
        case 5: launch_scan<5>(L, d, left, right, sp); break;
        case 6: launch_scan<6>(L, d, left, right, sp); break;
        case 7: launch_scan<7>(L, d, left, right, sp); break;
        case 8: launch_scan<8>(L, d, left, right, sp); break;
        case 9: launch_scan<9>(L, d, left, right, sp); break;
        case 10: launch_scan<10>(L, d, left, right, sp); break;
        case 11: launch_scan<11>(L, d, left, right, sp); break;
        case 12: launch_scan<12>(L, d, left, right, sp); break;
        case 13: launch_scan<13>(L, d, left, right, sp); break;
        default: launch_scan<16>(L, d, left, right, sp); break;  // Dn <= 512 (checked by the caller)
    }
}

}  // namespace tsm
