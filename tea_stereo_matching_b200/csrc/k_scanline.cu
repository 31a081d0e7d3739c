// k_scanline.cu -- four cascaded in-place scanline passes (scanlineOptimize,
// reference source/ADCensus.cpp:795-1011): down, up, right, left, per view, with the
// winner-take-all of cost2disparity (:1394-1413) fused into the last pass.
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
// bit-identical to the reference's sequential loop; min(a+P1, b+P1) == min(a,b)+P1
// because fp32 addition of a common term is monotonic.  All costs on this path are
// >= +0, so fp32 order equals the unsigned order of the bit patterns and the min over d
// is one REDUX.MIN.
//
// Mapping: one warp owns a line (a column for the vertical passes, a row for the
// horizontal ones) and walks it forward then backward; the predecessor's updated cost
// vector stays in registers (lanes over d, d = lane + 32k); the d-1 / d+1 neighbours
// come from two rotate-shuffles per register.  The other view's similarity bits for all
// K registers of a lane arrive as ONE 16-bit word (tflags, see k_prep.cu).  Inputs of the
// next pixels do not depend on the recurrence: they are loaded SCAN_PF steps ahead into a
// register ring.  The pass pair (down+up, right+left) is one launch; both views share it.
#include "tsm_common.cuh"
#include <limits.h>
#include <math_constants.h>

namespace tsm {

constexpr int SCAN_WARPS = 4;
constexpr int SCAN_PF = 4;           // prefetch distance
constexpr int SCAN_U = 2 * SCAN_PF;  // unroll factor = length of the register rings

struct ScanParams {
    float p1[3];
    float p2[3];
    int store_right_final;  // 0: the last pass of the right volume only feeds its WTA
};

struct StepIn {
    const float* src;      // main part of the pixel's cost vector (128-byte aligned)
    const float* tsrc;     // tail part (disparities >= Dm); read by the last register only
    const uint16_t* tf;    // other-view flag word of lane 0 for this pixel (lane l reads tf[sgn*l])
    const uint8_t* own;    // own-view flag byte
};

template <int K>
__device__ __forceinline__ void load_step(float (&cur)[K], unsigned& tf, unsigned& own, const StepIn& in, int lane, int sgn,
                                          int ownbit, bool has_tail, bool lastvalid)
{
#pragma unroll
    for (int k = 0; k < K; ++k) {
        if (k < K - 1 || !has_tail) cur[k] = in.src[lane + 32 * k];
        else cur[k] = lastvalid ? in.tsrc[lane] : CUDART_INF_F;
    }
    tf = in.tf[sgn * lane];
    own = (*in.own >> ownbit) & 1u;
}

// Updates prev (the predecessor's vector) to the new vector of this pixel; returns through
// `store` whether the pixel changed (m != 0).
template <int K>
__device__ __forceinline__ bool scan_step(float (&prev)[K], const float (&cur)[K], unsigned tf, unsigned own, int lane,
                                          const ScanParams& sp)
{
    unsigned mb = __float_as_uint(prev[0]);
#pragma unroll
    for (int k = 1; k < K; ++k) mb = min(mb, __float_as_uint(prev[k]));
    mb = __reduce_min_sync(0xffffffffu, mb);
    const float m = __uint_as_float(mb);
    if (m == 0.f) {  // ADCensus.cpp:880 -- pixel left untouched
#pragma unroll
        for (int k = 0; k < K; ++k) prev[k] = cur[k];
        return false;
    }
    // candidate penalties for "other view similar" = 0 / 1 (own is warp-uniform)
    const float p1a = own ? sp.p1[1] : sp.p1[0], p1b = own ? sp.p1[2] : sp.p1[1];
    const float mp2a = __fadd_rn(m, own ? sp.p2[1] : sp.p2[0]), mp2b = __fadd_rn(m, own ? sp.p2[2] : sp.p2[1]);
    float rl[K], rr[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        rl[k] = __shfl_sync(0xffffffffu, prev[k], (lane + 31) & 31);
        rr[k] = __shfl_sync(0xffffffffu, prev[k], (lane + 1) & 31);
    }
    const bool first = lane == 0, last = lane == 31;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const float lo = first ? (k > 0 ? rl[k > 0 ? k - 1 : 0] : CUDART_INF_F) : rl[k];
        const float hi = last ? (k < K - 1 ? rr[k < K - 1 ? k + 1 : K - 1] : CUDART_INF_F) : rr[k];
        const bool sim = (tf >> k) & 1u;
        const float p1 = sim ? p1b : p1a;
        const float mp2 = sim ? mp2b : mp2a;
        const float nb = __fadd_rn(fminf(lo, hi), p1);
        const float mo = fminf(fminf(mp2, prev[k]), nb);
        prev[k] = __fmul_rn(__fadd_rn(__fsub_rn(cur[k], m), mo), 0.5f);
    }
    return true;
}

template <int K>
__device__ __forceinline__ void store_vec(float* dst, float* tdst, const float (&v)[K], int lane, bool has_tail, bool lastvalid)
{
#pragma unroll
    for (int k = 0; k < K; ++k) {
        if (k < K - 1 || !has_tail) dst[lane + 32 * k] = v[k];
        else if (lastvalid) tdst[lane] = v[k];
    }
}

// cost2disparity: first strict minimum over d (ADCensus.cpp:1398-1409).
template <int K>
__device__ __forceinline__ int warp_argmin(const float (&v)[K], int lane)
{
    unsigned bb = __float_as_uint(v[0]);
    int bd = lane;
#pragma unroll
    for (int k = 1; k < K; ++k) {
        const unsigned b = __float_as_uint(v[k]);
        if (b < bb) { bb = b; bd = lane + 32 * k; }
    }
    const unsigned gmin = __reduce_min_sync(0xffffffffu, bb);
    return (int)__reduce_min_sync(0xffffffffu, bb == gmin ? (unsigned)bd : 0x7fffffffu);
}

// One direction of one line: `count` pixels starting at `first`, stepping by `dir` (+1/-1);
// the predecessor of a pixel is the previous one on the path.
//   VERT : pixel stride = W*Dp floats; flag row = row of max(pos, pred), flag bit 0
//   HORZ : pixel stride = Dp floats;   flag col = max(pos, pred),        flag bit 1
template <int K, bool VERT, bool WTA>
__device__ __forceinline__ void scan_dir(float (&prev)[K], const Vol& vol, const uint8_t* __restrict__ fown,
                                         const uint16_t* __restrict__ tfo, const Dims& dm, int line, int first, int dir,
                                         int count, int sgn, int lane, bool has_tail, bool lastvalid, bool do_store,
                                         int32_t* wta_out, const ScanParams& sp)
{
    const int W = dm.W, Wp = dm.W + 2 * kTfPad;
    const int ownbit = VERT ? 0 : 1;
    // element strides per step
    const ptrdiff_t vstep = (ptrdiff_t)dir * (VERT ? (ptrdiff_t)W * dm.Dm : (ptrdiff_t)dm.Dm);
    const ptrdiff_t wstep = (ptrdiff_t)dir * (VERT ? (ptrdiff_t)W * dm.Rp : (ptrdiff_t)dm.Rp);
    const ptrdiff_t fstep = (ptrdiff_t)dir * (VERT ? (ptrdiff_t)W : 1);
    const ptrdiff_t tstep = (ptrdiff_t)dir * (VERT ? (ptrdiff_t)Wp : 1);
    // position of the first pixel and of its flag pixel (= max(pos, pred))
    const int y0 = VERT ? first : line, x0 = VERT ? line : first;
    const int fy0 = VERT ? (dir > 0 ? first : first + 1) : line, fx0 = VERT ? line : (dir > 0 ? first : first + 1);
    StepIn nxt;  // operands of step i + SCAN_PF
    nxt.src = vol.main + ((size_t)y0 * W + x0) * dm.Dm;
    nxt.tsrc = vol.tail + ((size_t)y0 * W + x0) * dm.Rp;
    nxt.own = fown + (size_t)fy0 * W + fx0;
    nxt.tf = tfo + (size_t)(VERT ? 0 : 1) * dm.H * Wp + (size_t)fy0 * Wp + kTfPad + fx0;
    float* dst = vol.main + ((size_t)y0 * W + x0) * dm.Dm;
    float* tdst = vol.tail + ((size_t)y0 * W + x0) * dm.Rp;
    int32_t* wdst = WTA ? wta_out + (size_t)y0 * W + x0 : nullptr;

    // Register rings of 2*SCAN_PF entries: step i consumes entry i mod 2PF and refills entry
    // (i + PF) mod 2PF (consumed PF steps earlier), so a load never targets a live register.
    float cur[SCAN_U][K];
    unsigned tf[SCAN_U], own[SCAN_U];
#pragma unroll
    for (int u = 0; u < SCAN_U; ++u) {
        tf[u] = 0u;
        own[u] = 0u;
#pragma unroll
        for (int k = 0; k < K; ++k) cur[u][k] = 0.f;
    }
#pragma unroll
    for (int u = 0; u < SCAN_PF; ++u) {
        if (u < count) load_step<K>(cur[u], tf[u], own[u], nxt, lane, sgn, ownbit, has_tail, lastvalid);
        nxt.src += vstep;
        nxt.tsrc += wstep;
        nxt.own += fstep;
        nxt.tf += tstep;
    }
    for (int i0 = 0; i0 < count; i0 += SCAN_U) {
#pragma unroll
        for (int u = 0; u < SCAN_U; ++u) {
            const int i = i0 + u;
            if (i < count) {
                const int w = (u + SCAN_PF) % SCAN_U;
                if (i + SCAN_PF < count) load_step<K>(cur[w], tf[w], own[w], nxt, lane, sgn, ownbit, has_tail, lastvalid);
                nxt.src += vstep;
                nxt.tsrc += wstep;
                nxt.own += fstep;
                nxt.tf += tstep;
                const bool changed = scan_step<K>(prev, cur[u], tf[u], own[u], lane, sp);
                if (changed && do_store) store_vec<K>(dst, tdst, prev, lane, has_tail, lastvalid);
                dst += vstep;
                tdst += wstep;
                if (WTA) {
                    const int best = warp_argmin<K>(prev, lane);
                    if (lane == 0) *wdst = best;
                    wdst += fstep;
                }
            }
        }
    }
}

template <int K, bool VERT>
__global__ void __launch_bounds__(SCAN_WARPS * 32)
k_scanline(Dims dm, ViewPtrs v0, ViewPtrs v1, ScanParams sp, int32_t* wta0, int32_t* wta1)
{
    const int view = blockIdx.y;
    const ViewPtrs& v = view ? v1 : v0;
    const ViewPtrs& o = view ? v0 : v1;
    const int lane = threadIdx.x & 31;
    const int line = blockIdx.x * SCAN_WARPS + (threadIdx.x >> 5);
    const int nlines = VERT ? dm.W : dm.H, len = VERT ? dm.H : dm.W;
    if (line >= nlines) return;
    const int sgn = view == 0 ? 1 : -1;
    // K = ceil(Dn / 32) registers per lane; when Dn is not a multiple of 32 the last one holds the tail part
    const bool has_tail = dm.Rp != 0;
    const bool lastvalid = !has_tail || lane < dm.tail();
    float prev[K];
    {
        const size_t p0 = VERT ? (size_t)line : (size_t)line * dm.W;
        const float* src = v.vol.main + p0 * dm.Dm;
        const float* tsrc = v.vol.tail + p0 * dm.Rp;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            if (k < K - 1 || !has_tail) prev[k] = src[lane + 32 * k];
            else prev[k] = lastvalid ? tsrc[lane] : CUDART_INF_F;
        }
    }
    // forward: pos = 1 .. len-1 (pred pos-1); backward: pos = len-2 .. 0 (pred pos+1).
    scan_dir<K, VERT, false>(prev, v.vol, v.flags, o.tflags, dm, line, 1, 1, len - 1, sgn, lane, has_tail, lastvalid, true, nullptr, sp);
    if (VERT) {
        scan_dir<K, VERT, false>(prev, v.vol, v.flags, o.tflags, dm, line, len - 2, -1, len - 1, sgn, lane, has_tail, lastvalid,
                                 true, nullptr, sp);
    } else {
        // last pass: fuse the WTA; pixel len-1 is final after the forward pass.
        int32_t* wta_out = view ? wta1 : wta0;
        const int best = warp_argmin<K>(prev, lane);
        if (lane == 0) wta_out[(size_t)line * dm.W + len - 1] = best;
        const bool do_store = view == 0 || sp.store_right_final != 0;
        scan_dir<K, VERT, true>(prev, v.vol, v.flags, o.tflags, dm, line, len - 2, -1, len - 1, sgn, lane, has_tail, lastvalid,
                                do_store, wta_out, sp);
    }
}

template <int K>
static void launch_scan(const Launcher& L, const Dims& d, const ViewPtrs& left, const ViewPtrs& right, const ScanParams& sp,
                        int32_t* wta0, int32_t* wta1)
{
    dim3 gv((d.W + SCAN_WARPS - 1) / SCAN_WARPS, 2), gh((d.H + SCAN_WARPS - 1) / SCAN_WARPS, 2);
    k_scanline<K, true><<<gv, SCAN_WARPS * 32, 0, L.stream>>>(d, left, right, sp, wta0, wta1);
    k_scanline<K, false><<<gh, SCAN_WARPS * 32, 0, L.stream>>>(d, left, right, sp, wta0, wta1);
    L.count(2);
}

void scanline(const Launcher& L, const Dims& d, const ViewPtrs& left, const ViewPtrs& right, float p1_lo, float p2_lo,
              int32_t* wta_left, int32_t* wta_right, bool store_right_final)
{
    ScanParams sp;
    sp.p1[0] = p1_lo; sp.p1[1] = 0.25f; sp.p1[2] = 1.f;
    sp.p2[0] = p2_lo; sp.p2[1] = 0.75f; sp.p2[2] = 3.f;
    sp.store_right_final = store_right_final ? 1 : 0;
    const int K = (d.Dn + 31) / 32;
#define TSM_SCAN_CASE(k) case k: launch_scan<k>(L, d, left, right, sp, wta_left, wta_right); break;
    switch (K) {
        TSM_SCAN_CASE(1) TSM_SCAN_CASE(2) TSM_SCAN_CASE(3) TSM_SCAN_CASE(4) TSM_SCAN_CASE(5) TSM_SCAN_CASE(6)
        TSM_SCAN_CASE(7) TSM_SCAN_CASE(8) TSM_SCAN_CASE(9) TSM_SCAN_CASE(10) TSM_SCAN_CASE(11) TSM_SCAN_CASE(12)
        TSM_SCAN_CASE(13) TSM_SCAN_CASE(14) TSM_SCAN_CASE(15)
        default: launch_scan<16>(L, d, left, right, sp, wta_left, wta_right); break;  // Dn <= 512 (checked by the caller)
    }
#undef TSM_SCAN_CASE
}

}  // namespace tsm
