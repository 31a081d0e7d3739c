// k_scanline.cu -- four cascaded in-place scanline passes (scanlineOptimize,
// reference source/ADCensus.cpp:795-1011): down, up, right, left, per view, with the
// winner-take-all of cost2disparity (:1394-1413) fused into the last pass.
//
// This is the STRIPED walk (a lane holds d = lane + 32 k).  Since round 2 the usual disparity ranges (a main part of 32..256
// or 384 levels plus a tail: every BASELINE configuration) run through the blocked walk of k_scanline3.cu, which is bit-identical
// and 12 % faster; this file serves the other geometries (no tail, fewer than 32 levels, the 641-level ROI ranges), holds the
// recurrence / bit-exactness notes both walks rely on, and is the TSM_SCAN3=0 side of tests/test_gpu_scanline3.py.
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
// K registers of a lane, plus the own-view bit, arrive as ONE 32-bit word (scan table, see
// k_prep.cu).  The pass pair (down+up, right+left) is one launch; both views share it.
//
// Inputs of the next pixels do not depend on the recurrence, so they are streamed by TMA:
// lane 0 of every warp issues cp.async.bulk copies (the 128-byte aligned main part of the
// pixel's cost vector, its tail chunk and a 144-byte window of the scan table) NST steps
// ahead into a per-warp shared-memory ring; each stage completes on its own mbarrier
// (expect_tx bytes).  Completion is tracked by the mbarrier, not by register scoreboards,
// so the loads really stay in flight across the shuffle / REDUX waits of the recurrence
// (a register prefetch ring measured 4.2 + 3.5 ms per pair: every short-scoreboard wait also
// waited for the DRAM loads sharing its slot).
#include "tsm_common.cuh"
#include <cstdlib>
#include <limits.h>
#include <math_constants.h>

namespace tsm {

#ifndef TSM_SCAN_WARPS_H
#define TSM_SCAN_WARPS_H 1
#endif
#ifndef TSM_SCAN_WARPS
#define TSM_SCAN_WARPS 4
#endif
#ifndef TSM_SC_NST
#define TSM_SC_NST 0
#endif
// Warps (= lines) per CTA.  Vertical launch: 4.  Horizontal launch: 1 -- a 1080p pair has only 2160 rows for 148 SMs, and with
// single-warp CTAs the busiest SM holds 15 warps instead of 16 (the kernel is issue-bound: 2.64 -> 2.56 ms; vertical unchanged).
template <bool VERT>
struct ScanWarps {
    static constexpr int N = VERT ? TSM_SCAN_WARPS : TSM_SCAN_WARPS_H;
};
// TMA stages (steps in flight) per warp: 12 while a stage is small (K <= 8 registers = up to 256 disparities,
// <= 1.2 KB per stage: measured 5.47 -> 5.31 ms against 8 stages at K = 7; 16 stages cost a CTA per SM and lose),
// 8 for the wide vectors.  -DTSM_SC_NST=n overrides it for experiments.
template <int K>
struct ScanCfg {
    static constexpr int NST = TSM_SC_NST ? TSM_SC_NST : (K <= 8 ? 12 : 8);
};
constexpr int SC_WIN = 36;   // scan-table words fetched per step (32 lanes + 16-byte alignment slack)

struct ScanParams {
    float p1[3];
    float p2[3];
    int store_right_final;  // 0: the last pass of the right volume only feeds its WTA
    int stage_bytes;        // Dm*4 + tail chunk + table window, multiple of 16
    int tail_bytes;         // bytes of the tail chunk copied per step (0 when Dn % 32 == 0)
    int zero;               // always 0 (see the stage refill in scan_dir)
};

// ---- mbarrier / TMA bulk-copy primitives (sm_90+ PTX) ----
__device__ __forceinline__ void mbar_init(uint32_t bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void* src, unsigned bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// One elected lane of the (converged) warp: unlike `lane == 0` the compiler knows that exactly one thread is
// active inside, so the uniform-register bulk-copy instructions need no per-thread loop around them.
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

// (P1, P2) of the three similarity classes, held in registers for the whole walk
struct Penalties {
    float p1[3], p2[3];
};

// Updates prev (the predecessor's vector) to the new vector of this pixel; returns whether
// the pixel changed (m != 0).
template <int K>
__device__ __forceinline__ bool scan_step(float (&prev)[K], const float (&cur)[K], unsigned tf, unsigned own, int lane,
                                          const Penalties& pen, bool hole = false)
{
    if (hole) {  // mask matching: the predecessor is a masked pixel, the step is skipped (ADCensus.cpp:824, 862)
#pragma unroll
        for (int k = 0; k < K; ++k) prev[k] = cur[k];
        return false;
    }
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
    const float p1a = own ? pen.p1[1] : pen.p1[0], p1b = own ? pen.p1[2] : pen.p1[1];
    const float mp2a = __fadd_rn(m, own ? pen.p2[1] : pen.p2[0]), mp2b = __fadd_rn(m, own ? pen.p2[2] : pen.p2[1]);
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
// (With minD != 0 only the planes minD .. Dn - 1 take part: the caller then runs the stand-alone k_wta instead, k_post.cu.)
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

// Per-warp pipeline state: stage ring + mbarriers live in shared memory.
struct ScanPipe {
    uint32_t stage0;  // shared address of stage 0
    uint32_t bar0;    // shared address of mbarrier 0
    uint32_t slot;    // stage the consumer reads next
    uint32_t parity;  // phase parity the consumer waits for on that stage
};

// One direction of one line: `count` pixels starting at `first`, stepping by `dir` (+1/-1);
// the predecessor of a pixel is the previous one on the path.
//   VERT : pixel stride = one image row; flag pixel = (max(pos, pred), line), table plane 0
//   HORZ : pixel stride = one pixel;     flag pixel = (line, max(pos, pred)), table plane 1
// All per-step geometry is kept as running offsets (pixel index, table word offset, global pointers,
// stage address): the address arithmetic of the copy issue is lane-0-only code the whole warp waits for.
template <int K, bool VERT, bool WTA>
__device__ __forceinline__ void scan_dir(float (&prev)[K], const Vol& vol, const uint32_t* __restrict__ stab, const Dims& dm,
                                         ScanPipe& pipe, int line, int first, int dir, int count, int sgn, int lane,
                                         bool has_tail, bool lastvalid, bool do_store, int32_t* wta_out, const ScanParams& sp)
{
    constexpr int SC_NST = ScanCfg<K>::NST;
    const int W = dm.W, Wp = dm.stab_pitch();
    const uint32_t* tab = stab + (size_t)(VERT ? 0 : 1) * dm.H * Wp;
    const unsigned main_bytes = (unsigned)dm.Dm * 4u, tail_bytes = (unsigned)sp.tail_bytes;
    const unsigned total_bytes = main_bytes + tail_bytes + SC_WIN * 4u;
    const unsigned stage_bytes = (unsigned)sp.stage_bytes;
    const int lo_off = sgn > 0 ? 0 : -31;  // lowest table column a lane reads, relative to the flag column
    const bool wide_tail = dm.Rp >= 4;     // else the 16-byte tail chunk holds two pixels

    // step i: pixel index p(i) = p0 + i*pstep; its flag pixel sits at table word t(i) = t0 + i*tstep
    // (row*Wp + kTfPad + column; Wp % 4 == 0, so the 16-byte aligned window starts at (t + lo_off) & ~3)
    const int pstep = dir * (VERT ? W : 1), tstep = dir * (VERT ? Wp : 1);
    const int p0 = VERT ? first * W + line : line * W + first;
    const int f0 = dir > 0 ? first : first + 1;  // flag pixel = max(pos, pred) along the path
    const int t0 = (VERT ? f0 * Wp + line : line * Wp + f0) + kTfPad + lo_off;
    const ptrdiff_t vstep = (ptrdiff_t)pstep * dm.Dm, wstep = (ptrdiff_t)pstep * dm.Rp;

    // ---- producer (one elected lane issues; the running state is kept by every lane so the warp stays uniform) ----
    // The three copies of a step read from running 64-bit pointers wherever the source advances linearly along the path
    // (always for the main part; for the narrow tail chunk and the 16-byte aligned table window on vertical paths with an even
    // row pitch): recomputing them from the pixel / table index cost ~20 uniform instructions per step inside the elected
    // lane's block, which the whole warp waits for.
    const float* gmain = vol.main + (size_t)p0 * dm.Dm;
    int pi = p0, ti = t0;
    const bool tail_linear = wide_tail || (VERT && (W & 1) == 0);  // else (pi & ~1) does not advance by a constant
    const float* gtail = vol.tail + (wide_tail ? (size_t)p0 * dm.Rp : (size_t)(p0 & ~1) * 2);
    const ptrdiff_t tail_inc = wide_tail ? (ptrdiff_t)pstep * dm.Rp : (ptrdiff_t)pstep * 2;
    const uint32_t* gtab = tab + (t0 & ~3);  // vertical: Wp % 4 == 0, the aligned window start advances by tstep words
    uint32_t islot = pipe.slot;
    // arm the stage's mbarrier and launch the three bulk copies of the producer's current pixel
    auto issue = [&](uint32_t st, uint32_t bar) {
        mbar_expect_tx(bar, total_bytes);
        if (main_bytes) tma_load_1d(st, gmain, main_bytes, bar);
        if (tail_bytes) tma_load_1d(st + main_bytes, tail_linear ? gtail : vol.tail + (size_t)(pi & ~1) * 2, tail_bytes, bar);
        tma_load_1d(st + main_bytes + tail_bytes, VERT ? gtab : tab + (ti & ~3), SC_WIN * 4u, bar);
    };
    auto advance_producer = [&]() {
        gmain += vstep;
        gtail += tail_inc;
        if (VERT) gtab += tstep;
        pi += pstep;
        ti += tstep;
    };
    {
        const int npro = count < SC_NST ? count : SC_NST;
        for (int i = 0; i < npro; ++i) {
            if (elect_one()) issue(pipe.stage0 + islot * stage_bytes, pipe.bar0 + islot * 8u);
            advance_producer();
            islot = islot + 1 == SC_NST ? 0 : islot + 1;
        }
    }

    // ---- consumer ----
    float* dst = vol.main + (size_t)p0 * dm.Dm;
    float* tdst = vol.tail + (size_t)p0 * dm.Rp;
    int32_t* wdst = WTA ? wta_out + p0 : nullptr;
    int pc = p0, tc = t0;
    uint32_t st = pipe.stage0 + pipe.slot * stage_bytes, bar = pipe.bar0 + pipe.slot * 8u;
    // per-lane shared-memory offsets that do not change along a vertical path (the row pitches W and Wp are even / multiples
    // of 4 there): the tail element's slot in its 16-byte chunk and the table window index of the flag pixel
    // The narrow tail chunk (16 bytes) holds two pixels: the lane's element sits at float (pixel & 1) * 2 + lane, which
    // toggles along a path exactly when the pixel index changes parity with every step.
    const bool tail_slot_fixed = wide_tail || (VERT && (W & 1) == 0);
    uint32_t tail_off = main_bytes + ((wide_tail ? 0 : (p0 & 1) * 2) + lane) * 4;
    uint32_t tail_toggle = tail_slot_fixed ? 0u : 8u;
    const uint32_t win0 = main_bytes + tail_bytes;
    const int w0_fixed = (t0 & 3) - lo_off;
    uint32_t lane_win = (uint32_t)(sgn * lane * 4);  // byte offset of the lane's table word relative to the flag pixel's
    uint32_t win_fixed = win0 + (uint32_t)(w0_fixed * 4);  // vertical paths: the flag pixel's word in the window never moves
    asm volatile("" : "+r"(tail_off), "+r"(tail_toggle), "+r"(lane_win), "+r"(win_fixed));
    // The consumer's strides live in ordinary (per-thread) registers: everything here is warp-uniform, the compiler keeps
    // warp-uniform values in the 63 uniform registers, runs out of them in this loop and then RE-DERIVES the strides from the
    // kernel parameters in every iteration (64-bit multiplies: ~25 instructions per step in the first version).  An opaque
    // asm makes the copies "divergent" for the compiler.
    Penalties pen;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        pen.p1[c] = sp.p1[c];
        pen.p2[c] = sp.p2[c];
        asm volatile("" : "+f"(pen.p1[c]), "+f"(pen.p2[c]));
    }
    ptrdiff_t vstep_c = vstep, wstep_c = wstep;
    int pstep_c = pstep, tstep_c = tstep;
    asm volatile("" : "+l"(vstep_c), "+l"(wstep_c), "+r"(pstep_c), "+r"(tstep_c));

    for (int i = 0; i < count; ++i) {
        mbar_wait(bar, pipe.parity);
        float cur[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            if (k < K - 1 || !has_tail) cur[k] = lds_f32(st + (lane + 32 * k) * 4);
            else cur[k] = lastvalid ? lds_f32(st + tail_off) : CUDART_INF_F;
        }
        // shared address of the flag pixel's own table word
        const uint32_t wown = st + (VERT ? win_fixed : win0 + (uint32_t)(((tc & 3) - lo_off) * 4));
        const uint32_t tw = lds_u32(wown + lane_win);
        const uint32_t ow = lds_u32(wown);

        // Refill the stage.  The bulk copy writes shared memory through the async proxy and is NOT
        // ordered behind this warp's outstanding ld.shared: with a shared-memory-heavy kernel of
        // another stream co-resident on the SM (slow LDS) and the volume L2-resident (fast copy), the
        // copy overtook the loads and a step consumed the data of step i+8 (seen as rare run-to-run
        // differences with several contexts in flight).  Four ways to close it were built and measured at
        // C3 (ms per pair, both launches; profiles/README.md):
        //   TSM_SCAN_REFILL_DEP          5.47  round 1: copy address data-dependent on every loaded register
        //   (default) proxy fence         5.54  fence.proxy.async of the issuing lane after __syncwarp
        //   TSM_SCAN_REFILL_LATE         6.40  "empty" mbarrier (all lanes arrive, producer waits) AFTER the step used the values
        //   TSM_SCAN_REFILL_EMPTY_EARLY  6.44  the same right after the loads: NOT safe, see below
        // The empty-mbarrier phase parity equals that of the full barrier: both complete once per use of the stage.
#ifdef TSM_SCAN_REFILL_DEP
        // round-1 form, kept for A/B: the destination address is data-dependent on every loaded register
        uint32_t dep = tw ^ ow;
#pragma unroll
        for (int k = 0; k < K; ++k) dep ^= __float_as_uint(cur[k]);
        dep &= (uint32_t)sp.zero;
        __syncwarp();  // every lane has read the stage
        if (i + SC_NST < count && elect_one()) issue(st + dep, bar);
#elif defined(TSM_SCAN_REFILL_EMPTY_EARLY)
        // measured and REJECTED: the arrive issues while the ld.shared above are still in flight (nothing consumes their
        // results before it), the copy overtakes them and the determinism test fails again (250 pixels at D = 48)
        mbar_arrive(bar + SC_NST * 8u);
        if (i + SC_NST < count && elect_one()) {
            mbar_wait(bar + SC_NST * 8u, pipe.parity);
            issue(st, bar);
        }
#elif !defined(TSM_SCAN_REFILL_LATE)
        // default: the warp's generic-proxy reads of the stage (all lanes: __syncwarp) are ordered before the async-proxy
        // copy by a proxy fence of the issuing lane -- the PTX mechanism for exactly this pair of proxies
        __syncwarp();
        if (i + SC_NST < count && elect_one()) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue(st, bar);
        }
#endif
#ifndef TSM_SCAN_REFILL_LATE
        advance_producer();
#endif

        const bool changed = scan_step<K>(prev, cur, tw & 0x1fffffffu, ow >> 31, lane, pen, ((ow >> (dir > 0 ? 30 : 29)) & 1u) != 0);
        if (changed && do_store) store_vec<K>(dst, tdst, prev, lane, has_tail, lastvalid);
        dst += vstep_c;
        tdst += wstep_c;
        if (WTA) {
            const int best = warp_argmin<K>(prev, lane);
            if (lane == 0) *wdst = best;
            wdst += pstep_c;
        }
#ifdef TSM_SCAN_REFILL_LATE
        // release after the step has consumed the loaded values
        mbar_arrive(bar + SC_NST * 8u);
        if (i + SC_NST < count && elect_one()) {
            mbar_wait(bar + SC_NST * 8u, pipe.parity);
            issue(st, bar);
        }
        advance_producer();
#endif
        pc += pstep_c;
        tc += tstep_c;
        tail_off ^= tail_toggle;
        st += stage_bytes;
        bar += 8u;
        if (++pipe.slot == SC_NST) {
            pipe.slot = 0;
            pipe.parity ^= 1u;
            st = pipe.stage0;
            bar = pipe.bar0;
        }
    }
}

template <int K, bool VERT>
__global__ void __launch_bounds__(ScanWarps<VERT>::N * 32)
k_scanline(Dims dm, ViewPtrs v0, ViewPtrs v1, ScanParams sp, int32_t* wta0, int32_t* wta1)
{
    extern __shared__ __align__(128) unsigned char scan_smem[];
    const int view = blockIdx.y;
    const ViewPtrs& v = view ? v1 : v0;
    // warp index through a lane-0 broadcast: the compiler then knows it (and the line, the pointers and the stage
    // addresses derived from it) is warp-uniform and keeps the bulk-copy operands in uniform registers
    const int lane = threadIdx.x & 31, warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int line = blockIdx.x * ScanWarps<VERT>::N + warp;
    const int nlines = VERT ? dm.W : dm.H, len = VERT ? dm.H : dm.W;
    if (line >= nlines) return;
    const int sgn = view == 0 ? 1 : -1;
    // K = ceil(Dn / 32) registers per lane; when Dn is not a multiple of 32 the last one holds the tail part
    const bool has_tail = dm.Rp != 0;
    const bool lastvalid = !has_tail || lane < dm.tail();

    constexpr int SC_NST = ScanCfg<K>::NST;
    ScanPipe pipe;
    const uint32_t warp_bytes = SC_NST * (unsigned)sp.stage_bytes + 2 * SC_NST * 8u;  // stages, "full" barriers, "empty" barriers
    pipe.stage0 = (uint32_t)__cvta_generic_to_shared(scan_smem) + warp * ((warp_bytes + 127u) & ~127u);
    pipe.bar0 = pipe.stage0 + SC_NST * (unsigned)sp.stage_bytes;
    pipe.slot = 0;
    pipe.parity = 0;
    if (lane == 0) {
        for (int s = 0; s < SC_NST; ++s) {
            mbar_init(pipe.bar0 + s * 8u, 1);                 // full: one expect_tx arrival + the copies' bytes
            mbar_init(pipe.bar0 + (SC_NST + s) * 8u, 32);     // empty: every lane of the warp after its loads
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

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
    scan_dir<K, VERT, false>(prev, v.vol, v.stab, dm, pipe, line, 1, 1, len - 1, sgn, lane, has_tail, lastvalid, true, nullptr, sp);
    // The backward pass re-reads, through the async proxy (TMA), what the lanes of this warp have just
    // written with ordinary stores: make the stores visible at device scope, synchronise the warp, then
    // order them against the async proxy before lane 0 issues the first bulk copies.
    __threadfence();
    __syncwarp();
    asm volatile("fence.proxy.async;" ::: "memory");
    if (VERT) {
        scan_dir<K, VERT, false>(prev, v.vol, v.stab, dm, pipe, line, len - 2, -1, len - 1, sgn, lane, has_tail, lastvalid, true,
                                 nullptr, sp);
    } else {
        // last pass: fuse the WTA; pixel len-1 is final after the forward pass.
        int32_t* wta_out = view ? wta1 : wta0;
        const int best = warp_argmin<K>(prev, lane);
        if (lane == 0) wta_out[(size_t)line * dm.W + len - 1] = best;
        const bool do_store = view == 0 || sp.store_right_final != 0;
        scan_dir<K, VERT, true>(prev, v.vol, v.stab, dm, pipe, line, len - 2, -1, len - 1, sgn, lane, has_tail, lastvalid, do_store,
                                wta_out, sp);
    }
}

// =====================================================================================================================
// Two pixels per pipeline stage (k_scanline2).
//
// The walk above is issue-bound (72 % of the issue slots, ~235 warp instructions per step at K = 7), and more than half of a
// step is bookkeeping that does not depend on K: the mbarrier wait, the proxy fence and the three bulk-copy issues of the
// refill (one elected lane, the warp waits), the ring / pointer updates.  Two x-adjacent pixels are 2 * Dm * 4 contiguous
// bytes in the volume, share one 32-byte tail chunk and one 36-word table window, so ONE stage can carry both:
//   horizontal launch: the two pixels are consecutive steps of the row's path (pairs aligned to even x; the first / last pair
//                      of a path may have only one pixel on it);
//   vertical launch:   the two pixels are the same step of two neighbouring columns -- a warp walks TWO lines, two independent
//                      recurrences (twice the registers, twice the instruction-level parallelism).
// Per pixel the stage / fence / copy / loop overhead halves; the arithmetic (scan_step) is unchanged.
template <int K>
struct Scan2Cfg {
#ifndef TSM_SC2_NST
#define TSM_SC2_NST 0
#endif
    static constexpr int NST = TSM_SC2_NST ? TSM_SC2_NST : (K <= 8 ? 8 : 5);  // stages per warp, each two pixels deep
};

template <int K, bool VERT, bool WTA>
__device__ __forceinline__ void scan_dir2(float (&prev)[VERT ? 2 : 1][K], const Vol& vol, const uint32_t* __restrict__ stab, const Dims& dm,
                                          ScanPipe& pipe, int unit, int dir, int sgn, int lane, bool has_tail, bool lastvalid,
                                          bool do_store, int32_t* wta_out, const ScanParams& sp)
{
    constexpr int NST = Scan2Cfg<K>::NST;
    const int W = dm.W, H = dm.H, Wp = dm.stab_pitch();
    const uint32_t* tab = stab + (size_t)(VERT ? 0 : 1) * H * Wp;
    const unsigned main_bytes = (unsigned)dm.Dm * 4u;                      // one pixel
    const bool wide_tail = dm.Rp >= 4;
    const unsigned tail_bytes = !has_tail ? 0u : (wide_tail ? 2u * dm.Rp * 4u : 32u);  // two pixels (narrow: the 4-pixel chunk around them)
    const unsigned total_bytes = 2 * main_bytes + tail_bytes + SC_WIN * 4u;
    const unsigned stage_bytes = (unsigned)sp.stage_bytes;
    const int lo_off = sgn > 0 ? 0 : -31;

    // ---- geometry of the stages ----
    // pmin = pixel index of the stage's first (lower-x) pixel; tmin = table word of the lower of its two flag pixels
    int nstages, pmin, tmin, pstep, tstep;
    int lo = 0, hi = 0;  // HORZ: x range of the path
    if (VERT) {
        const int x0 = 2 * unit, first = dir > 0 ? 1 : H - 2, f0 = dir > 0 ? first : first + 1;
        nstages = H - 1;
        pmin = first * W + x0;
        tmin = f0 * Wp + kTfPad + x0;
        pstep = dir * W;
        tstep = dir * Wp;
    } else {
        lo = dir > 0 ? 1 : 0;
        hi = dir > 0 ? W - 1 : W - 2;
        const int m0 = dir > 0 ? lo / 2 : hi / 2;
        nstages = hi / 2 - lo / 2 + 1;
        pmin = unit * W + 2 * m0;
        // flag pixel of x = max(x, predecessor): forward x, backward x + 1 -> the lower one of the pair (2m, 2m+1) is 2m / 2m+1
        tmin = unit * Wp + kTfPad + 2 * m0 + (dir > 0 ? 0 : 1);
        pstep = dir * 2;
        tstep = dir * 2;
    }

    // ---- producer ----
    const float* gmain = vol.main + (size_t)pmin * dm.Dm;
    const ptrdiff_t vstep = (ptrdiff_t)pstep * dm.Dm;
    int ppi = pmin, pti = tmin;
    uint32_t islot = pipe.slot;
    auto issue = [&](uint32_t st, uint32_t bar) {
        mbar_expect_tx(bar, total_bytes);
        if (main_bytes) tma_load_1d(st, gmain, 2 * main_bytes, bar);
        if (tail_bytes)
            tma_load_1d(st + 2 * main_bytes, wide_tail ? vol.tail + (size_t)ppi * dm.Rp : vol.tail + (size_t)(ppi & ~1) * 2, tail_bytes, bar);
        tma_load_1d(st + 2 * main_bytes + tail_bytes, tab + ((pti + lo_off) & ~3), SC_WIN * 4u, bar);
    };
    auto advance_producer = [&]() {
        gmain += vstep;
        ppi += pstep;
        pti += tstep;
    };
    {
        const int npro = nstages < NST ? nstages : NST;
        for (int i = 0; i < npro; ++i) {
            if (elect_one()) issue(pipe.stage0 + islot * stage_bytes, pipe.bar0 + islot * 8u);
            advance_producer();
            islot = islot + 1 == NST ? 0 : islot + 1;
        }
    }

    // ---- consumer ----
    Penalties pen;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        pen.p1[c] = sp.p1[c];
        pen.p2[c] = sp.p2[c];
        asm volatile("" : "+f"(pen.p1[c]), "+f"(pen.p2[c]));
    }
    int pc = pmin, tc = tmin;  // of the stage being consumed
    int pstep_c = pstep, tstep_c = tstep;
    uint32_t lane_win = (uint32_t)(sgn * lane * 4), lane4 = (uint32_t)lane * 4u;
    asm volatile("" : "+r"(pstep_c), "+r"(tstep_c), "+r"(lane_win), "+r"(lane4));
    uint32_t st = pipe.stage0 + pipe.slot * stage_bytes, bar = pipe.bar0 + pipe.slot * 8u;
    const uint32_t win0 = 2 * main_bytes + tail_bytes;
    const int Rp = dm.Rp;
    // running per-lane output pointers of the stage's first pixel (the second one is one pixel vector further); everything the
    // stores need lives in per-thread registers (see scan_dir: the compiler otherwise re-derives it from the kernel parameters)
    char* dmain = reinterpret_cast<char*>(vol.main + (size_t)pmin * dm.Dm) + lane * 4;
    char* dtail = reinterpret_cast<char*>(vol.tail + (size_t)pmin * Rp) + lane * 4;
    int32_t* dwta = WTA ? wta_out + pmin : nullptr;
    ptrdiff_t dstep_m = (ptrdiff_t)pstep * dm.Dm * 4, dstep_t = (ptrdiff_t)pstep * Rp * 4;
    uint32_t mb = main_bytes, rp4 = (uint32_t)Rp * 4u;
    asm volatile("" : "+l"(dstep_m), "+l"(dstep_t), "+r"(mb), "+r"(rp4));
    const bool second_line = VERT && 2 * unit + 1 < W;

    for (int i = 0; i < nstages; ++i) {
        mbar_wait(bar, pipe.parity);
        // sub-step q handles the pixel in stage slot sl(q): path order = slot order except on a backward horizontal path
        float cur[2][K];
        uint32_t tw[2], ow[2];
        bool act[2];
        const int wbase = tc + lo_off - ((tc + lo_off) & ~3);  // window index of table word tc + lo_off
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int sl = (VERT || dir > 0) ? q : 1 - q;
            const int x = VERT ? 0 : (pc - unit * W) + sl;
            act[q] = VERT ? (sl == 0 || second_line) : (x >= lo && x <= hi);
            // table word of this pixel's flag pixel: tc is the lower one of the pair
            const int tw_idx = wbase - lo_off + sl;  // = (tc + sl) - window start
            const uint32_t wown = st + win0 + (uint32_t)(tw_idx * 4);
            tw[q] = lds_u32(wown + lane_win);
            ow[q] = lds_u32(wown);
            const uint32_t mbase = st + (uint32_t)sl * mb + lane4;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                if (k < K - 1 || !has_tail) cur[q][k] = lds_f32(mbase + 128u * k);
                else {
                    const int tslot = wide_tail ? sl * Rp : ((pc & 1) + sl) * 2;  // float index of the pixel's tail vector in the chunk
                    cur[q][k] = lastvalid ? lds_f32(st + 2 * main_bytes + (uint32_t)(tslot * 4) + lane4) : CUDART_INF_F;
                }
            }
        }
        // refill: the warp's generic-proxy reads of the stage are ordered before the async-proxy copy by a proxy fence of the issuing lane
        __syncwarp();
        if (i + NST < nstages && elect_one()) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue(st, bar);
        }
        advance_producer();
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            if (!act[q]) continue;  // warp-uniform
            const int sl = (VERT || dir > 0) ? q : 1 - q;
            float (&pv)[K] = prev[VERT ? q : 0];
            const bool changed = scan_step<K>(pv, cur[q], tw[q] & 0x1fffffffu, ow[q] >> 31, lane, pen,
                                              ((ow[q] >> (dir > 0 ? 30 : 29)) & 1u) != 0);
            if (changed && do_store) {
                char* om = dmain + (size_t)((uint32_t)sl * mb);
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    if (k < K - 1 || !has_tail) *reinterpret_cast<float*>(om + 128 * k) = pv[k];
                    else if (lastvalid) *reinterpret_cast<float*>(dtail + (size_t)((uint32_t)sl * rp4)) = pv[k];
                }
            }
            if (WTA) {
                const int best = warp_argmin<K>(pv, lane);
                if (lane == 0) dwta[sl] = best;
            }
        }
        dmain += dstep_m;
        dtail += dstep_t;
        if (WTA) dwta += pstep_c;
        pc += pstep_c;
        tc += tstep_c;
        st += stage_bytes;
        bar += 8u;
        if (++pipe.slot == NST) {
            pipe.slot = 0;
            pipe.parity ^= 1u;
            st = pipe.stage0;
            bar = pipe.bar0;
        }
    }
}

template <int K, bool VERT>
__global__ void __launch_bounds__(ScanWarps<VERT>::N * 32)
k_scanline2(Dims dm, ViewPtrs v0, ViewPtrs v1, ScanParams sp, int32_t* wta0, int32_t* wta1)
{
    extern __shared__ __align__(128) unsigned char scan_smem[];
    const int view = blockIdx.y;
    const ViewPtrs& v = view ? v1 : v0;
    const int lane = threadIdx.x & 31, warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int unit = blockIdx.x * ScanWarps<VERT>::N + warp;  // vertical: columns 2 unit, 2 unit + 1; horizontal: row
    const int nunits = VERT ? (dm.W + 1) / 2 : dm.H;
    if (unit >= nunits) return;
    const int sgn = view == 0 ? 1 : -1;
    const bool has_tail = dm.Rp != 0;
    const bool lastvalid = !has_tail || lane < dm.tail();
    constexpr int NST = Scan2Cfg<K>::NST;
    constexpr int NL = VERT ? 2 : 1;
    ScanPipe pipe;
    const uint32_t warp_bytes = NST * (unsigned)sp.stage_bytes + NST * 8u;
    pipe.stage0 = (uint32_t)__cvta_generic_to_shared(scan_smem) + warp * ((warp_bytes + 127u) & ~127u);
    pipe.bar0 = pipe.stage0 + NST * (unsigned)sp.stage_bytes;
    pipe.slot = 0;
    pipe.parity = 0;
    if (lane == 0) {
        for (int s = 0; s < NST; ++s) mbar_init(pipe.bar0 + s * 8u, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    float prev[NL][K];
    auto load_prev = [&](int l, size_t p0, bool valid) {
        const float* src = v.vol.main + p0 * dm.Dm;
        const float* tsrc = v.vol.tail + p0 * dm.Rp;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            if (!valid) prev[l][k] = CUDART_INF_F;
            else if (k < K - 1 || !has_tail) prev[l][k] = src[lane + 32 * k];
            else prev[l][k] = lastvalid ? tsrc[lane] : CUDART_INF_F;
        }
    };
    if (VERT) {
        load_prev(0, (size_t)2 * unit, true);
        load_prev(NL - 1, (size_t)2 * unit + 1, 2 * unit + 1 < dm.W);
    } else {
        load_prev(0, (size_t)unit * dm.W, true);
    }
    scan_dir2<K, VERT, false>(prev, v.vol, v.stab, dm, pipe, unit, 1, sgn, lane, has_tail, lastvalid, true, nullptr, sp);
    // the backward pass re-reads through the async proxy what this warp has just written with ordinary stores
    __threadfence();
    __syncwarp();
    asm volatile("fence.proxy.async;" ::: "memory");
    if (VERT) {
        scan_dir2<K, VERT, false>(prev, v.vol, v.stab, dm, pipe, unit, -1, sgn, lane, has_tail, lastvalid, true, nullptr, sp);
    } else {
        int32_t* wta_out = view ? wta1 : wta0;
        const int best = warp_argmin<K>(prev[0], lane);
        if (lane == 0) wta_out[(size_t)unit * dm.W + dm.W - 1] = best;
        const bool do_store = view == 0 || sp.store_right_final != 0;
        scan_dir2<K, VERT, true>(prev, v.vol, v.stab, dm, pipe, unit, -1, sgn, lane, has_tail, lastvalid, do_store, wta_out, sp);
    }
}

static bool scan2_enabled()
{
    static const bool on = [] {
        const char* e = getenv("TSM_SCAN2");
        return !(e && e[0] == '0');
    }();
    return on;
}

template <int K>
static void launch_scan_vertical(const Launcher& L, const Dims& d, const ViewPtrs& left, const ViewPtrs& right, const ScanParams& sp,
                                 int32_t* wta0, int32_t* wta1);

template <int K>
static void launch_scan2(const Launcher& L, const Dims& d, const ViewPtrs& left, const ViewPtrs& right, ScanParams sp,
                         int32_t* wta0, int32_t* wta1)
{
    const ScanParams sp1 = sp;  // one-pixel stage geometry (vertical launch)
    constexpr int NST = Scan2Cfg<K>::NST;
    const unsigned tail2 = d.Rp == 0 ? 0u : (d.Rp >= 4 ? 2u * d.Rp * 4u : 32u);
    sp.stage_bytes = 2 * d.Dm * 4 + (int)tail2 + SC_WIN * 4;
    const unsigned warp_bytes = ((unsigned)(NST * sp.stage_bytes + NST * 8) + 127u) & ~127u;
    constexpr int WV = ScanWarps<true>::N, WH = ScanWarps<false>::N;
    const size_t smem_v = (size_t)WV * warp_bytes, smem_h = (size_t)WH * warp_bytes, smem = smem_v > smem_h ? smem_v : smem_h;
    static PerDevice smem_set;
    if (smem > smem_set.cur()) {
        cudaFuncSetAttribute(k_scanline2<K, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_scanline2<K, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        smem_set.cur() = smem;
    }
    const int nunits_v = (d.W + 1) / 2;
    dim3 gv((nunits_v + WV - 1) / WV, 2), gh((d.H + WH - 1) / WH, 2);
    // Measured at C3 (ms): vertical launch 2.41 with one pixel per stage, 2.45 with two (half as many warps, each with two
    // recurrences: the issue slots are not filled better); horizontal launch 2.58 -> 2.48.  So only the horizontal launch uses it.
    static const bool vert2 = getenv("TSM_SCAN2_VERT") != nullptr;
    if (!vert2) {
        launch_scan_vertical<K>(L, d, left, right, sp1, wta0, wta1);
    } else {
        L.begin("scanline/vertical");
        k_scanline2<K, true><<<gv, WV * 32, smem_v, L.stream>>>(d, left, right, sp, wta0, wta1);
        L.end();
        L.count(1);
    }
    L.begin("scanline/horizontal");
    k_scanline2<K, false><<<gh, WH * 32, smem_h, L.stream>>>(d, left, right, sp, wta0, wta1);
    L.end();
    L.count(1);
}

template <int K>
static size_t scan1_smem(const ScanParams& sp, bool vert)
{
    constexpr int SC_NST = ScanCfg<K>::NST;
    const unsigned warp_bytes = ((unsigned)(SC_NST * sp.stage_bytes + 2 * SC_NST * 8) + 127u) & ~127u;
    constexpr int WV = ScanWarps<true>::N, WH = ScanWarps<false>::N;
    const size_t smem_v = (size_t)WV * warp_bytes, smem_h = (size_t)WH * warp_bytes, smem = smem_v > smem_h ? smem_v : smem_h;
    static PerDevice smem_set;
    if (smem > smem_set.cur()) {
        cudaFuncSetAttribute(k_scanline<K, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_scanline<K, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        smem_set.cur() = smem;
    }
    return vert ? smem_v : smem_h;
}

template <int K>
static void launch_scan_vertical(const Launcher& L, const Dims& d, const ViewPtrs& left, const ViewPtrs& right, const ScanParams& sp,
                                 int32_t* wta0, int32_t* wta1)
{
    constexpr int WV = ScanWarps<true>::N;
    const size_t smem_v = scan1_smem<K>(sp, true);
    dim3 gv((d.W + WV - 1) / WV, 2);
    L.begin("scanline/vertical");
    k_scanline<K, true><<<gv, WV * 32, smem_v, L.stream>>>(d, left, right, sp, wta0, wta1);
    L.end();
    L.count(1);
}

template <int K>
static void launch_scan(const Launcher& L, const Dims& d, const ViewPtrs& left, const ViewPtrs& right, const ScanParams& sp,
                        int32_t* wta0, int32_t* wta1)
{
    constexpr int WH = ScanWarps<false>::N;
    launch_scan_vertical<K>(L, d, left, right, sp, wta0, wta1);
    const size_t smem_h = scan1_smem<K>(sp, false);
    dim3 gh((d.H + WH - 1) / WH, 2);
    L.begin("scanline/horizontal");
    k_scanline<K, false><<<gh, WH * 32, smem_h, L.stream>>>(d, left, right, sp, wta0, wta1);
    L.end();
    L.count(1);
}

void scanline(const Launcher& L, const Dims& d, const ViewPtrs& left, const ViewPtrs& right, float p1_lo, float p2_lo,
              int32_t* wta_left, int32_t* wta_right, bool store_right_final)
{
    ScanParams sp;
    sp.p1[0] = p1_lo; sp.p1[1] = 0.25f; sp.p1[2] = 1.f;
    sp.p2[0] = p2_lo; sp.p2[1] = 0.75f; sp.p2[2] = 3.f;
    sp.store_right_final = store_right_final ? 1 : 0;
    sp.zero = 0;
    sp.tail_bytes = d.Rp == 0 ? 0 : (d.Rp >= 4 ? d.Rp * 4 : 16);
    sp.stage_bytes = d.Dm * 4 + sp.tail_bytes + SC_WIN * 4;
    const int K = (d.Dn + 31) / 32;
    // two pixels per stage for the usual disparity ranges (K <= 13: up to 416 levels); the one-pixel walk for the wide ones
#define TSM_SCAN2_CASE(k) case k: launch_scan2<k>(L, d, left, right, sp, wta_left, wta_right); return;
    if (scan2_enabled() && d.W >= 4 && d.H >= 2) {
        switch (K) {
            TSM_SCAN2_CASE(1) TSM_SCAN2_CASE(2) TSM_SCAN2_CASE(3) TSM_SCAN2_CASE(4) TSM_SCAN2_CASE(5) TSM_SCAN2_CASE(6)
            TSM_SCAN2_CASE(7) TSM_SCAN2_CASE(8) TSM_SCAN2_CASE(9) TSM_SCAN2_CASE(10) TSM_SCAN2_CASE(11) TSM_SCAN2_CASE(12)
            TSM_SCAN2_CASE(13)
            default: break;
        }
    }
#undef TSM_SCAN2_CASE
#define TSM_SCAN_CASE(k) case k: launch_scan<k>(L, d, left, right, sp, wta_left, wta_right); break;
    switch (K) {
        TSM_SCAN_CASE(1) TSM_SCAN_CASE(2) TSM_SCAN_CASE(3) TSM_SCAN_CASE(4) TSM_SCAN_CASE(5) TSM_SCAN_CASE(6)
        TSM_SCAN_CASE(7) TSM_SCAN_CASE(8) TSM_SCAN_CASE(9) TSM_SCAN_CASE(10) TSM_SCAN_CASE(11) TSM_SCAN_CASE(12)
        TSM_SCAN_CASE(13) TSM_SCAN_CASE(14) TSM_SCAN_CASE(15) TSM_SCAN_CASE(16)
        // ROI / mask matching searches W / 2 + 1 levels (ADCensus.cpp:339-340): 641 at the reference's 1280-wide demo size
        TSM_SCAN_CASE(17) TSM_SCAN_CASE(18) TSM_SCAN_CASE(19) TSM_SCAN_CASE(20) TSM_SCAN_CASE(21) TSM_SCAN_CASE(22)
        TSM_SCAN_CASE(23)
        default: launch_scan<24>(L, d, left, right, sp, wta_left, wta_right); break;  // Dn <= kMaxLevels (checked by the caller)
    }
#undef TSM_SCAN_CASE
}

}  // namespace tsm
