// k_scanline3.cu -- the four scanline passes (scanlineOptimize, reference source/ADCensus.cpp:795-1011) as a BLOCKED walk.
// Same arithmetic, same pass order and same in-place volume as k_scanline.cu (see its header for the recurrence and the
// bit-exactness argument); what changes is who holds which disparity, where the flags come from and who issues the copies.
//
// k_scanline.cu is issue-bound: ~225 warp instructions per pixel step at D = 192, of which only ~75 are the recurrence.
//   * There a lane holds d = lane + 32 k, so BOTH neighbours d-1 / d+1 of every register come from another lane: two
//     rotate-shuffles and two edge selects per register.  Here a lane holds KM = Dm / 32 CONSECUTIVE disparities of the main
//     part (d = KM * lane + k), the neighbours are its own registers, and a step needs four shuffles in all (the two block
//     ends and the tail register, which stays striped: d = Dm + lane).
//   * The cost vector of a pixel is contiguous in d, so a blocked lane reads its KM values with vector loads straight from
//     the TMA stage (lane stride KM * 4 bytes: conflict-free for the vector width used) and writes them back with vector stores
//     (3 x STG.64 at a 24-byte lane stride measured 6.16 TB/s against 6.24 TB/s for fully coalesced stores: scripts/micro/blk.cu).
//   * The similarity flags of the other view are needed for KM consecutive columns per lane: they come from bit planes
//     (one bit per pixel, SbLayout in tsm_common.cuh) with one funnel shift, instead of one table word per (pixel, lane).
//     A vertical step stages a 48-byte window of the flag row (was 144 bytes), a horizontal path keeps its whole flag row in
//     shared memory.  The pixel's own flag and the mask-matching "black predecessor" bits are one nibble per pixel of a
//     string along the path, also in shared memory.
//   * Vertical launch (k_scan3v): a CTA owns neighbouring columns, a stage is one image row of them -- three bulk copies per
//     step for all its lines, issued by a PRODUCER warp (one elected lane) that owns the address arithmetic; the consumer
//     warps wait on the stage's "full" mbarrier and release it through its "empty" mbarrier.
//     Horizontal launch (k_scan3h): a stage is a group of consecutive pixels of a row, refilled by the warp itself.
//   * The refill hazard of k_scanline.cu (async-proxy copy overtaking a warp's outstanding ld.shared) is settled by program
//     order: every loaded register is consumed by arithmetic that the step's stores wait for, and the stores precede the
//     release / refill (see step3).
// ~160 warp instructions per step at D = 192; 4.97 -> 4.36 ms per 1080p pair (0.86 / 0.82 of the HBM copy bandwidth).
#include "tsm_common.cuh"
#include <cstdlib>
#include <limits.h>
#include <math_constants.h>

namespace tsm {
namespace {

template <int KM>
struct S3Cfg {
    static constexpr int V = (KM % 4 == 0) ? 4 : (KM % 2 == 0) ? 2 : 1;  // floats per vector access
};

struct S3Params {
    float p1[3], p2[3];
    int store_right_final;  // 0: the last pass of the right volume only feeds its WTA
    int nw;                 // consumer warps (= lines) per CTA; one more warp produces
    int nst;                // stages of a ring
    int P;                  // horizontal launch: pixels per stage (even)
    int main_all;           // bytes of the cost vectors of a stage (nw or P pixels)
    int tail_bytes;         // bytes of the tail chunk of a stage
    int win_words;          // vertical launch: words of the flag-row window of a stage
    int stage_bytes;        // main_all + tail_bytes (+ win_words * 4), multiple of 16
    int warp_bytes;         // horizontal launch: shared memory of one warp (ring, barriers, strings), multiple of 128
    SbLayout lay;
};

// ---- mbarrier / TMA bulk-copy primitives (sm_90+ PTX), as in k_scanline.cu ----
__device__ __forceinline__ void mbar_init(uint32_t bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_lane0(uint32_t bar, int lane)
{
    asm volatile("{ .reg .pred p; setp.eq.s32 p, %1, 0; @p mbarrier.arrive.release.cta.shared::cta.b64 _, [%0]; }" ::"r"(bar), "r"(lane)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
// the producer's wait: it is ahead of its consumers most of the time, so it must not burn issue slots polling
#ifndef TSM_S3_SLEEP
#define TSM_S3_SLEEP 200
#endif
__device__ __forceinline__ void mbar_wait_sleepy(uint32_t bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "WAIT_LOOP:\n"
        "nanosleep.u32 %2;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(bar), "r"(parity), "n"(TSM_S3_SLEEP) : "memory");
}
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void* src, unsigned bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{ .reg .pred P; elect.sync _|P, 0xffffffff; selp.u32 %0, 1, 0, P; }" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ float lds_f32(uint32_t a)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t a)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v)
{
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}

// the lane's KM consecutive floats, from shared / global memory and back
template <int KM>
__device__ __forceinline__ void lds_block(float (&v)[KM], uint32_t a)
{
    constexpr int V = S3Cfg<KM>::V;
#pragma unroll
    for (int q = 0; q < KM / V; ++q) {
        if (V == 4)
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(v[4 * q]), "=f"(v[4 * q + 1]), "=f"(v[4 * q + 2]), "=f"(v[4 * q + 3]) : "r"(a + 16 * q) : "memory");
        else if (V == 2)
            asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v[2 * q]), "=f"(v[2 * q + 1]) : "r"(a + 8 * q) : "memory");
        else
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v[q]) : "r"(a + 4 * q) : "memory");
    }
}
template <int KM>
__device__ __forceinline__ void ldg_block(float (&v)[KM], const float* p)
{
    constexpr int V = S3Cfg<KM>::V;
#pragma unroll
    for (int q = 0; q < KM / V; ++q) {
        if (V == 4) {
            const float4 t = reinterpret_cast<const float4*>(p)[q];
            v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
        } else if (V == 2) {
            const float2 t = reinterpret_cast<const float2*>(p)[q];
            v[2 * q] = t.x; v[2 * q + 1] = t.y;
        } else {
            v[q] = p[q];
        }
    }
}
template <int KM>
__device__ __forceinline__ void stg_block(float* p, const float (&v)[KM])
{
    constexpr int V = S3Cfg<KM>::V;
#pragma unroll
    for (int q = 0; q < KM / V; ++q) {
        if (V == 4) reinterpret_cast<float4*>(p)[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        else if (V == 2) reinterpret_cast<float2*>(p)[q] = make_float2(v[2 * q], v[2 * q + 1]);
        else p[q] = v[q];
    }
}

struct Pen3 {
    float p1[3], p2[3];
};

// One step of partialOptimization (ADCensus.cpp:869-913) on the blocked vector: pm[k] is disparity KM * lane + k, pt the
// tail element Dm + lane (+inf in the lanes beyond the tail).  sim: bit k = the other view's pair is similar for pm[k];
// simt the same for the tail element.  Returns whether the pixel changed.
// One step of partialOptimization (ADCensus.cpp:869-913) on the blocked vector: pm[k] is disparity KM * lane + k, pt the
// tail element Dm + lane (+inf in the lanes beyond the tail).  sim: bit k = the other view's pair is similar for pm[k];
// simt the same for the tail element.
// A skipped step (black predecessor in mask matching, :824 / :862; m == 0, :880) leaves the pixel as it is.  It is computed
// and selected away instead of branched around, and the caller stores the vector in either case: every loaded register is
// then consumed by arithmetic that the stores wait for, and the stores precede the release / refill of the stage in program
// order -- the bulk copy into the stage can never be issued while a ld.shared of it is still in flight (the hazard of
// k_scanline.cu), without a per-step dependency chain on the release.  Re-storing an unchanged vector costs nothing extra:
// the traffic budget counts every cell as written.
template <int KM>
__device__ __forceinline__ void step3(float (&pm)[KM], float& pt, const float (&cm)[KM], float ct, unsigned sim, bool simt,
                                      unsigned own, bool hole, int lane, const Pen3& pen)
{
    unsigned mb = __float_as_uint(pt);
#pragma unroll
    for (int k = 0; k < KM; ++k) mb = min(mb, __float_as_uint(pm[k]));
    mb = __reduce_min_sync(0xffffffffu, mb);
    const float m = __uint_as_float(mb);
    const bool skip = hole || m == 0.f;
    const float p1a = own ? pen.p1[1] : pen.p1[0], p1b = own ? pen.p1[2] : pen.p1[1];
    const float mp2a = __fadd_rn(m, own ? pen.p2[1] : pen.p2[0]), mp2b = __fadd_rn(m, own ? pen.p2[2] : pen.p2[1]);
    // the four values that cross a lane boundary: block ends and the tail register
    const float A = __shfl_sync(0xffffffffu, pm[KM - 1], (lane + 31) & 31);  // d - 1 of pm[0]; in lane 0: d - 1 of the first tail element
    const float B = __shfl_sync(0xffffffffu, pm[0], (lane + 1) & 31);        // d + 1 of pm[KM - 1]
    const float C = __shfl_sync(0xffffffffu, pt, (lane + 31) & 31);          // d - 1 of the tail element
    const float D = __shfl_sync(0xffffffffu, pt, (lane + 1) & 31);           // d + 1 of the tail element; in lane 31: d + 1 of pm[KM - 1]
    const bool first = lane == 0, last = lane == 31;
    const float lo0 = first ? CUDART_INF_F : A, hiL = last ? D : B;
    const float lot = first ? A : C, hit = last ? CUDART_INF_F : D;
    float nm[KM];
#pragma unroll
    for (int k = 0; k < KM; ++k) {
        const float lo = k == 0 ? lo0 : pm[k > 0 ? k - 1 : 0];
        const float hi = k == KM - 1 ? hiL : pm[k < KM - 1 ? k + 1 : KM - 1];
        const bool s = (sim >> k) & 1u;
        const float p1 = s ? p1b : p1a;
        const float mp2 = s ? mp2b : mp2a;
        const float nb = __fadd_rn(fminf(lo, hi), p1);
        const float mo = fminf(fminf(mp2, pm[k]), nb);
        nm[k] = __fmul_rn(__fadd_rn(__fsub_rn(cm[k], m), mo), 0.5f);
    }
    {
        const float p1 = simt ? p1b : p1a;
        const float mp2 = simt ? mp2b : mp2a;
        const float nb = __fadd_rn(fminf(lot, hit), p1);
        const float mo = fminf(fminf(mp2, pt), nb);
        const float nt = __fmul_rn(__fadd_rn(__fsub_rn(ct, m), mo), 0.5f);
        pt = skip ? ct : nt;
    }
#pragma unroll
    for (int k = 0; k < KM; ++k) pm[k] = skip ? cm[k] : nm[k];
}

// cost2disparity: first strict minimum over d (ADCensus.cpp:1398-1409) of the blocked vector
template <int KM>
__device__ __forceinline__ int argmin3(const float (&pm)[KM], float pt, int lane, int Dm)
{
    unsigned bb = __float_as_uint(pm[0]);
    int bd = KM * lane;
#pragma unroll
    for (int k = 1; k < KM; ++k) {
        const unsigned b = __float_as_uint(pm[k]);
        if (b < bb) { bb = b; bd = KM * lane + k; }
    }
    const unsigned bt = __float_as_uint(pt);
    const unsigned gmin = __reduce_min_sync(0xffffffffu, min(bb, bt));
    const unsigned cand = bb == gmin ? (unsigned)bd : (bt == gmin ? (unsigned)(Dm + lane) : 0x7fffffffu);
    return (int)__reduce_min_sync(0xffffffffu, cand);
}

// ---- vertical launch: a CTA owns NW neighbouring columns; a stage is one image row of them -------------------------
// The NW cost vectors of a row are contiguous in the volume, so one bulk copy brings all of them, one more their tail
// elements and a third the flag-row window they share: three copies per step of NW lines.  The consumer warps share the
// stage ring (they wait on the same "full" barrier, each releases the stage once); the ring depth absorbs their jitter.
template <int KM>
__device__ __forceinline__ void walk_v(float (&pm)[KM], float& pt, const Vol& vol, const Dims& dm, const S3Params& sp, uint32_t base,
                                       uint32_t& slot, uint32_t& parity, int w, int line0, int first, int dir, int count, int sgn,
                                       int lane)
{
    const uint32_t NST = (uint32_t)sp.nst;
    const int W = dm.W, r = dm.tail(), line = line0 + w;
    const bool lastvalid = lane < r;
    const unsigned main_bytes = (unsigned)dm.Dm * 4u, tail_bytes = (unsigned)sp.tail_bytes, stage_bytes = (unsigned)sp.stage_bytes;
    const unsigned main_all = (unsigned)sp.main_all;
    const bool wide_tail = dm.Rp >= 4;  // else two pixels share a 16-byte chunk of the tail volume
    const uint32_t bars = base + NST * stage_bytes;                                      // full[NST], empty[NST], pass
    const uint32_t own_base = bars + 2 * NST * 8u + 16u + (uint32_t)(w * sp.lay.h8p) * 4u;  // the column's nibble string

    const int pstep = dir * W;
    const int p0 = first * W + line, pl0 = first * W + line0;
    int f = dir > 0 ? first : first + 1;  // flag pixel = max(pos, pred) along the path

    // the lane's flag bits: KM consecutive columns starting (s > 0) or ending (s < 0) at x + s * (minD + KM * lane); the tail
    // element's single column x + s * (minD + Dm + lane); all relative to the window start of the CTA's first column
    const int tl = lastvalid ? lane : 0;
    const int lowc = sgn > 0 ? dm.minD + KM * lane : -dm.minD - KM * lane - (KM - 1);
    const int tailc = sgn * (dm.minD + dm.Dm + tl);
    const int wsc = sgn > 0 ? line0 + dm.minD : line0 - dm.minD - (dm.Dn - 1);
    const int WS = (kSbPad + wsc) & ~127;
    const int rel = kSbPad + line + lowc - WS, trel = kSbPad + line + tailc - WS;
    uint32_t woff = main_all + tail_bytes + (uint32_t)(rel >> 5) * 4u, sh = (uint32_t)rel & 31u;
    uint32_t twoff = main_all + tail_bytes + (uint32_t)(trel >> 5) * 4u, tsh = (uint32_t)trel & 31u;
    uint32_t tail_off = main_all + (wide_tail ? 0u : (uint32_t)(pl0 & 1) * 8u) + (uint32_t)(w * dm.Rp + lane) * 4u;
    // odd width and two pixels per 16-byte chunk: the chunk offset of the row's first pixel alternates 0, 8, 0, ... along the path
    int tail_delta = (!wide_tail && (W & 1)) ? ((pl0 & 1) ? -8 : 8) : 0;
    uint32_t lane_off = (uint32_t)w * main_bytes + (uint32_t)(KM * 4 * lane);
    asm volatile("" : "+r"(tail_off), "+r"(tail_delta), "+r"(lane_off), "+r"(woff), "+r"(sh), "+r"(twoff), "+r"(tsh));

    Pen3 pen;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        pen.p1[c] = sp.p1[c];
        pen.p2[c] = sp.p2[c];
        asm volatile("" : "+f"(pen.p1[c]), "+f"(pen.p2[c]));
    }
    float* dst = vol.main + (size_t)p0 * dm.Dm + KM * lane;
    float* tdst = vol.tail + (size_t)p0 * dm.Rp + lane;
    ptrdiff_t vstep = (ptrdiff_t)pstep * dm.Dm, wstep = (ptrdiff_t)pstep * dm.Rp;
    int dir_c = dir;
    asm volatile("" : "+l"(vstep), "+l"(wstep), "+r"(dir_c));
    const unsigned hole_bit = dir > 0 ? 2u : 4u;

    uint32_t st = base + slot * stage_bytes, bar = bars + slot * 8u;
    for (int i = 0; i < count; ++i) {
        mbar_wait(bar, parity);
        float cm[KM];
        lds_block<KM>(cm, st + lane_off);
        float ct = lds_f32(st + tail_off);
        ct = lastvalid ? ct : CUDART_INF_F;
        const uint32_t w0 = lds_u32(st + woff), w1 = lds_u32(st + woff + 4u), wt = lds_u32(st + twoff);
        const uint32_t nw = lds_u32(own_base + (((uint32_t)f >> 3) << 2));
        uint32_t g = __funnelshift_r(w0, w1, sh);
        if (sgn < 0) g = __brev(g) >> (32 - KM);
        const bool simt = (wt >> tsh) & 1u;
        const uint32_t nib = nw >> (((uint32_t)f & 7u) * 4u);

        step3<KM>(pm, pt, cm, ct, g, simt, nib & 1u, (nib & hole_bit) != 0, lane, pen);
        stg_block<KM>(dst, pm);
        if (lastvalid) *tdst = pt;
        dst += vstep;
        tdst += wstep;
        // Release the stage to the producer: one arrival per warp, after the stores of every lane (see step3).
        __syncwarp();
        mbar_arrive_lane0(bar + NST * 8u, lane);
        f += dir_c;
        tail_off += (uint32_t)tail_delta;
        tail_delta = -tail_delta;
        st += stage_bytes;
        bar += 8u;
        if (++slot == NST) {
            slot = 0;
            parity ^= 1u;
            st = base;
            bar = bars;
        }
    }
}

__device__ __forceinline__ void produce_v(const Dims& dm, const S3Params& sp, const Vol& vol, const uint32_t* __restrict__ other_bits,
                                          uint32_t base, int line0, int nact, int sgn)
{
    const uint32_t NST = (uint32_t)sp.nst;
    const int W = dm.W, len = dm.H;
    const unsigned main_bytes = (unsigned)dm.Dm * 4u, tail_bytes = (unsigned)sp.tail_bytes, stage_bytes = (unsigned)sp.stage_bytes;
    const unsigned main_all = (unsigned)sp.main_all, win_bytes = (unsigned)sp.win_words * 4u;
    const unsigned copy_main = (unsigned)nact * main_bytes, total = copy_main + tail_bytes + win_bytes;
    const uint32_t bars = base + NST * stage_bytes, pass_bar = bars + 2 * NST * 8u;
    const int wsc = sgn > 0 ? line0 + dm.minD : line0 - dm.minD - (dm.Dn - 1);
    const int wsw = ((kSbPad + wsc) & ~127) >> 5;  // word offset of the CTA's flag window inside a flag row
    const char* const main_b = reinterpret_cast<const char*>(vol.main);
    const char* const tail_b = reinterpret_cast<const char*>(vol.tail);
    const unsigned tail_pitch = (unsigned)dm.Rp * 4u;
    uint32_t slot = 0, eparity = 1;  // a fresh "empty" barrier passes a wait on parity 1
    for (int pass = 0; pass < 2; ++pass) {
        const int dir = pass ? -1 : 1, first = pass ? len - 2 : 1, count = len - 1;
        if (pass) {
            // the backward pass re-reads, through the async proxy, what the consumers have just written with ordinary stores
            mbar_wait(pass_bar, 0);
            asm volatile("fence.proxy.async;" ::: "memory");
        }
        const int pstep = dir * W, fstep = dir * sp.lay.pitch;
        int p = first * W + line0;
        int frow = (dir > 0 ? first : first + 1) * sp.lay.pitch + wsw;
        uint32_t st = base, full = bars;
        st += slot * stage_bytes;
        full += slot * 8u;
        for (int i = 0; i < count; ++i) {
            mbar_wait_sleepy(full + NST * 8u, eparity);
            if (elect_one()) {
                mbar_expect_tx(full, total);
                tma_load_1d(st, main_b + (unsigned long long)(unsigned)p * main_bytes, copy_main, full);
                tma_load_1d(st + main_all, tail_b + (((unsigned long long)(unsigned)p * tail_pitch) & ~15ull), tail_bytes, full);
                tma_load_1d(st + main_all + tail_bytes, other_bits + (unsigned)frow, win_bytes, full);
            }
            __syncwarp();
            p += pstep;
            frow += fstep;
            st += stage_bytes;
            full += 8u;
            if (++slot == NST) {
                slot = 0;
                eparity ^= 1u;
                st = base;
                full = bars;
            }
        }
    }
}

template <int KM>
__global__ void __launch_bounds__(1024) k_scan3v(Dims dm, ViewPtrs v0, ViewPtrs v1, S3Params sp)
{
    extern __shared__ __align__(128) unsigned char s3_smem[];
    const int NW = sp.nw, NST = sp.nst;
    const int view = blockIdx.y;
    const ViewPtrs& v = view ? v1 : v0;
    const ViewPtrs& vo = view ? v0 : v1;
    const int lane = threadIdx.x & 31, warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int len = dm.H, line0 = blockIdx.x * NW;
    const int nact = min(NW, dm.W - line0);
    const int sgn = view == 0 ? 1 : -1;
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(s3_smem);
    const uint32_t bars = base + (uint32_t)NST * (uint32_t)sp.stage_bytes, pass_bar = bars + 2 * NST * 8u;
    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; ++s) {
            mbar_init(bars + s * 8u, 1);                       // full: one expect_tx arrival + the copies' bytes
            mbar_init(bars + (NST + s) * 8u, (unsigned)nact);  // empty: one arrival per consumer warp
        }
        mbar_init(pass_bar, (unsigned)nact);                   // every consumer warp, after the forward pass
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp == NW) {
        produce_v(dm, sp, v.vol, vo.sbits + sp.lay.bp_v, base, line0, nact, sgn);
        return;
    }
    if (warp >= nact) return;
    const int line = line0 + warp;
    {
        const uint32_t own_base = pass_bar + 16u + (uint32_t)(warp * sp.lay.h8p) * 4u;
        const uint32_t* src = v.sbits + sp.lay.own_v + (size_t)line * sp.lay.h8p;
        for (int i = lane; i < sp.lay.h8p; i += 32) sts_u32(own_base + 4u * i, src[i]);
        __syncwarp();
    }
    const bool lastvalid = lane < dm.tail();
    float pm[KM], pt;
    ldg_block<KM>(pm, v.vol.main + (size_t)line * dm.Dm + KM * lane);
    pt = lastvalid ? v.vol.tail[(size_t)line * dm.Rp + lane] : CUDART_INF_F;
    uint32_t slot = 0, parity = 0;
    // down: pos = 1 .. H-1 (pred pos-1); up: pos = H-2 .. 0 (pred pos+1).
    walk_v<KM>(pm, pt, v.vol, dm, sp, base, slot, parity, warp, line0, 1, 1, len - 1, sgn, lane);
    // hand the written column to the producer: stores visible at device scope and ordered against the async proxy
    __threadfence();
    asm volatile("fence.proxy.async;" ::: "memory");
    __syncwarp();
    mbar_arrive_lane0(pass_bar, lane);
    walk_v<KM>(pm, pt, v.vol, dm, sp, base, slot, parity, warp, line0, len - 2, -1, len - 1, sgn, lane);
}

// ---- horizontal launch: a warp owns a row; a stage is a group of P consecutive pixels of it ----------------------------------
// P pixels are contiguous in the volume: two bulk copies per group (cost vectors, tail elements) and one mbarrier wait per P
// steps.  Rows are far apart in memory, so nothing can be shared between warps: every warp refills its own ring (one
// elected lane, after the group has been consumed; a producer warp serving 15 rows was the bottleneck here: 2.34 ms against
// 2.15 ms with twice the group size).  Groups are aligned to multiples of P (P even: the 16-byte tail chunk then starts at
// the same offset in every group of a row).  The row's flag bits of the other view and its own nibble string stay in shared
// memory for the whole path.
template <int KM, bool WTA>
__device__ __forceinline__ void walk_h(float (&pm)[KM], float& pt, const Vol& vol, const Dims& dm, const S3Params& sp, uint32_t wbase,
                                       uint32_t& slot, uint32_t& parity, int line, int dir, int sgn, int lane, bool do_store,
                                       int32_t* wta_out)
{
    const uint32_t NST = (uint32_t)sp.nst;
    const int W = dm.W, r = dm.tail(), P = sp.P;
    const bool lastvalid = lane < r;
    const unsigned main_bytes = (unsigned)dm.Dm * 4u, stage_bytes = (unsigned)sp.stage_bytes, main_all = (unsigned)sp.main_all;
    const unsigned tail_bytes = (unsigned)sp.tail_bytes, tail_pitch = (unsigned)dm.Rp * 4u;
    const bool wide_tail = dm.Rp >= 4;
    const uint32_t bars = wbase + NST * stage_bytes;
    const uint32_t own_base = bars + NST * 8u;
    const uint32_t row_bits = own_base + (uint32_t)sp.lay.w8p * 4u;

    const int rowp = line * W;
    int x = dir > 0 ? 1 : W - 2;
    int g = x / P;
    const int ngroups = dir > 0 ? (W + P - 1) / P - g : g + 1;

    // ---- refills (one elected lane) ----
    const char* const main_b = reinterpret_cast<const char*>(vol.main);
    const char* const tail_b = reinterpret_cast<const char*>(vol.tail);
    auto issue = [&](int gg, uint32_t st, uint32_t bar) {
        const int gx = gg * P;
        const unsigned pg = (unsigned)(rowp + gx);
        const unsigned copy_main = (unsigned)min(P, W - gx) * main_bytes;
        mbar_expect_tx(bar, copy_main + tail_bytes);
        tma_load_1d(st, main_b + (unsigned long long)pg * main_bytes, copy_main, bar);
        tma_load_1d(st + main_all, tail_b + (((unsigned long long)pg * tail_pitch) & ~15ull), tail_bytes, bar);
    };
    {
        const int npro = ngroups < (int)NST ? ngroups : (int)NST;
        uint32_t is = slot;
        for (int k = 0; k < npro; ++k) {
            if (elect_one()) issue(g + dir * k, wbase + is * stage_bytes, bars + is * 8u);
            __syncwarp();
            is = is + 1 == NST ? 0 : is + 1;
        }
    }

    int f = dir > 0 ? x : x + 1;
    const int tl = lastvalid ? lane : 0;
    const int lowc = sgn > 0 ? dm.minD + KM * lane : -dm.minD - KM * lane - (KM - 1);
    const int tailc = sgn * (dm.minD + dm.Dm + tl);
    int Bl = kSbPad + f + lowc, Bt = kSbPad + f + tailc;  // running bit indices into the flag row
    uint32_t toff0 = main_all + (wide_tail ? 0u : (uint32_t)(rowp & 1) * 8u) + (uint32_t)lane * 4u;
    uint32_t lane_off = (uint32_t)(KM * 4 * lane);
    uint32_t tstride = tail_pitch, mstride = main_bytes;
    asm volatile("" : "+r"(toff0), "+r"(lane_off), "+r"(tstride), "+r"(mstride));

    Pen3 pen;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        pen.p1[c] = sp.p1[c];
        pen.p2[c] = sp.p2[c];
        asm volatile("" : "+f"(pen.p1[c]), "+f"(pen.p2[c]));
    }
    float* dst = vol.main + (size_t)(rowp + x) * dm.Dm + KM * lane;
    float* tdst = vol.tail + (size_t)(rowp + x) * dm.Rp + lane;
    int32_t* wdst = WTA ? wta_out + rowp + x : nullptr;
    ptrdiff_t vstep = (ptrdiff_t)dir * dm.Dm, wstep = (ptrdiff_t)dir * dm.Rp;
    int dir_c = dir;
    asm volatile("" : "+l"(vstep), "+l"(wstep), "+r"(dir_c));
    const unsigned hole_bit = dir > 0 ? 2u : 4u;

    uint32_t st = wbase + slot * stage_bytes, bar = bars + slot * 8u;
    for (int gi = 0; gi < ngroups; ++gi) {
        const int gx = g * P;
        const int xe = dir > 0 ? min(gx + P, W) - 1 : gx;  // last pixel of the group in walking order
        const int n = dir > 0 ? xe - x + 1 : x - xe + 1;
        uint32_t qm = st + (uint32_t)(x - gx) * mstride + lane_off;
        uint32_t qt = st + toff0 + (uint32_t)(x - gx) * tstride;
        mbar_wait(bar, parity);
        for (int j = 0; j < n; ++j) {
            float cm[KM];
            lds_block<KM>(cm, qm);
            float ct = lds_f32(qt);
            ct = lastvalid ? ct : CUDART_INF_F;
            const uint32_t a = row_bits + (((uint32_t)Bl >> 5) << 2);
            const uint32_t w0 = lds_u32(a), w1 = lds_u32(a + 4u);
            const uint32_t wt = lds_u32(row_bits + (((uint32_t)Bt >> 5) << 2));
            const uint32_t nw = lds_u32(own_base + (((uint32_t)f >> 3) << 2));
            uint32_t gb = __funnelshift_r(w0, w1, (uint32_t)Bl);
            if (sgn < 0) gb = __brev(gb) >> (32 - KM);
            const bool simt = (wt >> ((uint32_t)Bt & 31u)) & 1u;
            const uint32_t nib = nw >> (((uint32_t)f & 7u) * 4u);

            step3<KM>(pm, pt, cm, ct, gb, simt, nib & 1u, (nib & hole_bit) != 0, lane, pen);
            if (do_store) {
                stg_block<KM>(dst, pm);
                if (lastvalid) *tdst = pt;
            }
            dst += vstep;
            tdst += wstep;
            if (WTA) {
                const int best = argmin3<KM>(pm, pt, lane, dm.Dm);
                if (lane == 0) *wdst = best;
                wdst += dir_c;
            }
            f += dir_c;
            Bl += dir_c;
            Bt += dir_c;
            qm += (uint32_t)dir_c * mstride;
            qt += (uint32_t)dir_c * tstride;
        }
        x += dir_c * n;
        // Refill the stage with the group NST ahead.  The bulk copy writes shared memory through the async proxy and is not
        // ordered behind this warp's ld.shared of the stage by itself (k_scanline.cu: the copy once overtook loads that were
        // still in flight).  Every lane's loads of the group have been consumed by arithmetic that its stores (the cost
        // vectors, or the WTA word where the vectors are not kept) wait for, the stores precede this point in program order,
        // and __syncwarp collects the lanes (see step3).  -DTSM_S3_FENCE adds the proxy fence of the issuing lane in front
        // (no measurable difference).
        __syncwarp();
        if (gi + (int)NST < ngroups && elect_one()) {
#ifdef TSM_S3_FENCE
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
            issue(g + dir_c * (int)NST, st, bar);
        }
        g += dir_c;
        st += stage_bytes;
        bar += 8u;
        if (++slot == NST) {
            slot = 0;
            parity ^= 1u;
            st = wbase;
            bar = bars;
        }
    }
}

template <int KM>
__global__ void __launch_bounds__(512) k_scan3h(Dims dm, ViewPtrs v0, ViewPtrs v1, S3Params sp, int32_t* wta0, int32_t* wta1)
{
    extern __shared__ __align__(128) unsigned char s3_smem[];
    const int NW = sp.nw, NST = sp.nst;
    const int view = blockIdx.y;
    const ViewPtrs& v = view ? v1 : v0;
    const ViewPtrs& vo = view ? v0 : v1;
    const int lane = threadIdx.x & 31, warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int line = blockIdx.x * NW + warp;
    if (line >= dm.H) return;
    const int sgn = view == 0 ? 1 : -1;
    const uint32_t wbase = (uint32_t)__cvta_generic_to_shared(s3_smem) + warp * (uint32_t)sp.warp_bytes;
    const uint32_t bars = wbase + NST * (uint32_t)sp.stage_bytes;
    if (lane == 0) {
        for (int s = 0; s < NST; ++s) mbar_init(bars + s * 8u, 1);  // one expect_tx arrival + the copies' bytes
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {
        const uint32_t own_base = bars + NST * 8u;
        const uint32_t* src = v.sbits + sp.lay.own_h + (size_t)line * sp.lay.w8p;
        for (int i = lane; i < sp.lay.w8p; i += 32) sts_u32(own_base + 4u * i, src[i]);
        const uint32_t row_bits = own_base + (uint32_t)sp.lay.w8p * 4u;
        const uint32_t* rsrc = vo.sbits + sp.lay.bp_h + (size_t)line * sp.lay.pitch;
        for (int i = lane; i < sp.lay.pitch; i += 32) sts_u32(row_bits + 4u * i, rsrc[i]);
    }
    __syncwarp();
    const bool lastvalid = lane < dm.tail();
    float pm[KM], pt;
    const size_t p0 = (size_t)line * dm.W;
    ldg_block<KM>(pm, v.vol.main + p0 * dm.Dm + KM * lane);
    pt = lastvalid ? v.vol.tail[p0 * dm.Rp + lane] : CUDART_INF_F;
    uint32_t slot = 0, parity = 0;
    // right: x = 1 .. W-1 (pred x-1); left: x = W-2 .. 0 (pred x+1) with the WTA fused (pixel W-1 is final after the first pass)
    walk_h<KM, false>(pm, pt, v.vol, dm, sp, wbase, slot, parity, line, 1, sgn, lane, true, nullptr);
    // The second pass re-reads, through the async proxy, what the lanes of this warp have just written with ordinary stores:
    // make them visible at device scope, collect the warp, then order them against the async proxy before the first copies.
    __threadfence();
    __syncwarp();
    asm volatile("fence.proxy.async;" ::: "memory");
    int32_t* wta_out = view ? wta1 : wta0;
    const int best = argmin3<KM>(pm, pt, lane, dm.Dm);
    if (lane == 0) wta_out[p0 + dm.W - 1] = best;
    const bool do_store = view == 0 || sp.store_right_final != 0;
    walk_h<KM, true>(pm, pt, v.vol, dm, sp, wbase, slot, parity, line, -1, sgn, lane, do_store, wta_out);
}

// ---- flag bit planes and path strings ------------------------------------------------------------------------------
// flags[p]: bit 0 similar to the pixel above, bit 1 similar to the pixel to the left, bit 2 black (k_prep.cu, k_arms_flags)
__global__ void __launch_bounds__(256) k_scan_bits(const uint8_t* __restrict__ flags_left, const uint8_t* __restrict__ flags_right,
                                                   uint32_t* __restrict__ out_left, uint32_t* __restrict__ out_right, int H, int W, SbLayout lay)
{
    const uint8_t* __restrict__ f = blockIdx.y ? flags_right : flags_left;
    uint32_t* __restrict__ out = blockIdx.y ? out_right : out_left;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= lay.words) return;
    uint32_t w = 0;
    if (i < lay.own_h) {
        const int plane = i >= lay.bp_h ? 1 : 0;
        const size_t j = i - (plane ? lay.bp_h : 0);
        const int y = (int)(j / lay.pitch), wi = (int)(j % lay.pitch);
        const int c0 = wi * 32 - kSbPad;
        if (c0 + 31 >= 0 && c0 < W) {
#pragma unroll 8
            for (int b = 0; b < 32; ++b) {
                const int c = c0 + b;
                if (c >= 0 && c < W) w |= (uint32_t)((f[(size_t)y * W + c] >> plane) & 1u) << b;
            }
        }
    } else if (i < lay.own_v) {
        const size_t j = i - lay.own_h;
        const int y = (int)(j / lay.w8p), wi = (int)(j % lay.w8p);
        for (int n = 0; n < 8; ++n) {
            const int c = 8 * wi + n;
            if (c < W) {
                const unsigned fl = f[(size_t)y * W + c];
                unsigned nib = ((fl >> 1) & 1u) | (((fl >> 2) & 1u) << 2);
                if (c > 0) nib |= ((f[(size_t)y * W + c - 1] >> 2) & 1u) << 1;
                w |= nib << (4 * n);
            }
        }
    } else {
        const size_t j = i - lay.own_v;
        const int x = (int)(j / lay.h8p), wi = (int)(j % lay.h8p);
        for (int n = 0; n < 8; ++n) {
            const int y = 8 * wi + n;
            if (y < H) {
                const unsigned fl = f[(size_t)y * W + x];
                unsigned nib = (fl & 1u) | (((fl >> 2) & 1u) << 2);
                if (y > 0) nib |= ((f[(size_t)(y - 1) * W + x] >> 2) & 1u) << 1;
                w |= nib << (4 * n);
            }
        }
    }
    out[i] = w;
}

static int env_int(const char* name, int dflt)
{
    const char* e = getenv(name);
    return e && *e ? atoi(e) : dflt;
}

template <int KM>
void launch3(const Launcher& L, const Dims& d, const ViewPtrs& left, const ViewPtrs& right, S3Params sp, int32_t* wta0, int32_t* wta1)
{
    int dev = 0, nsm = 148;
    cudaGetDevice(&dev);
    static PerDevice sm_count;
    if (!sm_count.cur()) {
        cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
        sm_count.cur() = (size_t)nsm;
    }
    nsm = (int)sm_count.cur();
    static const int e_nwv = env_int("TSM_S3_NWV", 0), e_nstv = env_int("TSM_S3_NSTV", 0), e_nwh = env_int("TSM_S3_NWH", 0),
                     e_p = env_int("TSM_S3_P", 0), e_nsth = env_int("TSM_S3_NSTH", 0);
    static PerDevice smem_set_v, smem_set_h;
    const int main_bytes = d.Dm * 4;
    {
        // Vertical launch.  All columns of both views in ONE wave of two CTAs per SM: a column is a chain of H dependent steps, a
        // second, partial wave would run at the latency of that chain with most of the machine idle.
        S3Params s = sp;
        int nw = e_nwv ? e_nwv : (d.W + nsm - 1) / nsm;
        nw = nw < 4 ? 4 : (nw > 30 ? 30 : nw);
        const int budget = env_int("TSM_S3_BUDGET", 110) * 1024;
        int nst = 0;
        for (;; --nw) {
            s.main_all = nw * main_bytes;
            s.tail_bytes = d.Rp >= 4 ? nw * d.Rp * 4 : (8 + nw * 8 + 15) & ~15;
            s.win_words = ((((126 + d.Dn + nw - 1) >> 5) + 2) + 3) & ~3;
            s.stage_bytes = s.main_all + s.tail_bytes + s.win_words * 4;
            nst = (budget - nw * s.lay.h8p * 4 - 512) / s.stage_bytes;
            if (nst >= 4 || nw == 1) break;
        }
        if (e_nstv) nst = e_nstv;
        nst = nst > 16 ? 16 : (nst < 2 ? 2 : nst);
        s.nw = nw;
        s.nst = nst;
        const size_t smem = (size_t)nst * s.stage_bytes + 2 * nst * 8 + 16 + (size_t)nw * s.lay.h8p * 4;
        if (smem > smem_set_v.cur()) {
            cudaFuncSetAttribute(k_scan3v<KM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            smem_set_v.cur() = smem;
        }
        dim3 g((d.W + nw - 1) / nw, 2);
        L.begin("scanline/vertical");
        k_scan3v<KM><<<g, (nw + 1) * 32, smem, L.stream>>>(d, left, right, s);
        L.end();
    }
    {
        // Horizontal launch.  2 H independent rows, each a chain of 2 W dependent steps: the rows are dealt out in whole waves
        // of ONE CTA per SM (1080p on 148 SMs: 144 CTAs of 15 rows, measured 2.25 ms against 2.39 ms with one row per CTA,
        // where the block scheduler decides which SMs and which of their four warp schedulers get the 15th warp).
        S3Params s = sp;
        const int P = e_p ? (e_p & ~1) : (KM <= 8 ? 8 : 4);
        const int nst = e_nsth ? e_nsth : (KM <= 8 ? 2 : 3);
        s.P = P;
        s.main_all = P * main_bytes;
        s.tail_bytes = d.Rp >= 4 ? P * d.Rp * 4 : (8 + P * 8 + 15) & ~15;
        s.win_words = 0;
        s.stage_bytes = s.main_all + s.tail_bytes;
        s.nst = nst;
        s.warp_bytes = (nst * s.stage_bytes + nst * 8 + (s.lay.w8p + s.lay.pitch) * 4 + 127) & ~127;
        int nw = e_nwh;
        for (int waves = 1; nw == 0; ++waves) {
            const int n = (2 * d.H + nsm * waves - 1) / (nsm * waves);
            if (n <= 16 && (size_t)n * s.warp_bytes <= 220 * 1024) nw = n;
        }
        s.nw = nw;
        const size_t smem = (size_t)nw * s.warp_bytes;
        if (smem > smem_set_h.cur()) {
            cudaFuncSetAttribute(k_scan3h<KM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            smem_set_h.cur() = smem;
        }
        dim3 g((d.H + nw - 1) / nw, 2);
        L.begin("scanline/horizontal");
        k_scan3h<KM><<<g, nw * 32, smem, L.stream>>>(d, left, right, s, wta0, wta1);
        L.end();
    }
    L.count(2);
}

}  // namespace

bool scanline3_geometry(const Dims& d)
{
    if (d.Rp == 0 || d.W < 4 || d.H < 2) return false;
    const int km = d.Dm / 32;
    return (km >= 1 && km <= 8) || km == 12;
}

bool scanline3_supported(const Dims& d)
{
    // TSM_SCAN3=0 selects k_scanline.cu for every geometry (read per call: the parity test switches it inside one process)
    const char* e = getenv("TSM_SCAN3");
    return !(e && e[0] == '0') && scanline3_geometry(d);
}

void prep_scan_bits(const Launcher& L, const Dims& d, const uint8_t* flags_left, const uint8_t* flags_right, uint32_t* sbits_left,
                    uint32_t* sbits_right)
{
    const SbLayout lay = sb_layout(d.H, d.W);
    dim3 g((unsigned)((lay.words + 255) / 256), 2);
    k_scan_bits<<<g, 256, 0, L.stream>>>(flags_left, flags_right, sbits_left, sbits_right, d.H, d.W, lay);
    L.count(1);
}

void scanline3(const Launcher& L, const Dims& d, const ViewPtrs& left, const ViewPtrs& right, float p1_lo, float p2_lo,
               int32_t* wta_left, int32_t* wta_right, bool store_right_final)
{
    S3Params sp;
    sp.p1[0] = p1_lo; sp.p1[1] = 0.25f; sp.p1[2] = 1.f;
    sp.p2[0] = p2_lo; sp.p2[1] = 0.75f; sp.p2[2] = 3.f;
    sp.store_right_final = store_right_final ? 1 : 0;
    sp.nw = sp.nst = sp.P = sp.main_all = sp.tail_bytes = sp.win_words = sp.stage_bytes = sp.warp_bytes = 0;
    sp.lay = sb_layout(d.H, d.W);
    switch (d.Dm / 32) {
#define TSM_S3_CASE(k) case k: launch3<k>(L, d, left, right, sp, wta_left, wta_right); break;
        TSM_S3_CASE(1) TSM_S3_CASE(2) TSM_S3_CASE(3) TSM_S3_CASE(4) TSM_S3_CASE(5) TSM_S3_CASE(6) TSM_S3_CASE(7) TSM_S3_CASE(8)
        TSM_S3_CASE(12)
#undef TSM_S3_CASE
        default: break;  // scanline3_supported() said no
    }
}

}  // namespace tsm
