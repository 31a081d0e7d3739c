// k_aggregate.cu -- cross-based cost aggregation (costAggregate, reference
// source/ADCensus.cpp:685-793), 4 iterations, pass order H,V | V,H | H,V | V,H.
//
// One aggregation1D pass is  out(p) = sum_{j=-a(p)}^{b(p)} in(p + j*r)  along a line
// (ADCensus.cpp:711-718); after the second pass of an iteration the plane is divided
// by the cross-window size (:743-749).  The reference adds in fp32 in ascending order;
// here every (line, d) chain walks its line once, keeps a running fp64 prefix sum P and
// emits  P[o+b+1] - P[o-a]  with a lag of 33 (= max arm); the last 72 prefixes live in a
// ring in shared memory or in tensor memory.  fp64 prefix differences stay within ~4e-7
// relative of the reference's sequential fp32 sums after all 4 iterations (fp32 prefixes do
// not).  Each cell is read once and written once per pass, in place.
//
// Thread mapping (walk_line): one thread owns TWO adjacent disparities (d, d+1) of one
// line: 8-byte global accesses, 16-byte ring entries, and the per-step index arithmetic
// (ring slots, arm decode, pointers) is paid once per two cells.  Lanes run over d, so a
// warp reads 256 contiguous bytes per step.  Inputs do not depend on the recurrence: they
// are loaded in batches of AGG_PF steps into NB register buffers, NB-1 batches ahead of
// their use, so the only loop-carried dependency is the fp64 prefix add.  A batch is
// processed in sub-blocks (pushes, ring loads, outputs) so that the ring round trip is
// paid once per four steps.  The main loop is unrolled NB*AGG_PF times and the ring length
// is a multiple of AGG_PF, so the ring slots of the pushes are compile-time offsets.
//
// Kernels: k_agg_persist (rings in shared AND tensor memory, one persistent CTA per SM, warps
// pull work items from a global counter; needs Dm % 64 == 0 = every BASELINE configuration),
// k_agg_walk (shared-memory rings only, any Dm), k_agg_small (lines shorter than 66),
// k_agg_tail_h (horizontal passes of the tail part: staged row + block scan).
//
// Division: the reference divides the fp32 sum by (float)N with an IEEE fp32 divide
// (ADCensus.cpp:747).  The walk multiplies by the correctly rounded reciprocal stored next to the
// step descriptors and applies one residual correction (div_exact_rn, tsm_common.cuh); the small
// kernels refine rcp.approx themselves (div_exact).  Both equal __fdiv_rn for every N <= 4489
// (tsm_selftest), without its slow-path branch.
//
// Per-position side data (the two arm lengths of the pass and N) come as ONE 32-bit step
// descriptor (k_prep.cu), four consecutive positions per 16-byte load: small broadcast loads
// cost real bandwidth on this access pattern (measured: -15 % with two scalar loads per step).
#include "tsm_common.cuh"
#include <cstdio>
#include <cstdlib>

namespace tsm {

#ifndef TSM_AGG_NC
#define TSM_AGG_NC 2
#endif
constexpr int AGG_NC = TSM_AGG_NC;        // adjacent disparities (chains) per thread: 1 or 2
constexpr int AGG_BLOCK = 128 / AGG_NC;   // threads per CTA
constexpr int AGG_LAG = kMaxArm;          // 33
#ifndef TSM_AGG_PF
#define TSM_AGG_PF 4
#endif
constexpr int AGG_PF = TSM_AGG_PF;        // prefetch distance
// Register buffers of the walk: NB-1 batches of AGG_PF loads in flight, NB*AGG_PF steps unrolled per loop
// iteration.  More buffers = deeper prefetch but a bigger loop body, and the instruction cache matters: with
// two ring homes (two copies of the loop) resident per SM, 8 buffers spent more cycles in no_instruction
// stalls than in memory stalls (ncu), 6 was the measured optimum without the L2 prefetch below, 4 with it; the single-copy
// kernel is best at 8.
#ifndef TSM_AGG_NBUF
#define TSM_AGG_NBUF 4
#endif
constexpr int AGG_NBUF_PERSIST = TSM_AGG_NBUF, AGG_NBUF_STATIC = 8;
constexpr int AGG_U = 8 * AGG_PF;         // upper bound of the steps per main-loop iteration (over-read slack, minimum line length)
#ifndef TSM_AGG_SUB
#define TSM_AGG_SUB 4
#endif
constexpr int AGG_SUB = TSM_AGG_SUB;      // steps per software-pipelined sub-block of a batch
static_assert(AGG_PF % AGG_SUB == 0, "sub-blocks tile a batch");
constexpr int AGG_RING = 72;              // >= 68 prefixes, multiple of AGG_PF
// P[i] lives in slot (i + AGG_P0) mod AGG_RING, chosen so that P[AGG_LAG + 1] (the first
// prefix written in the main loop) sits on a multiple of AGG_PF.
constexpr int AGG_P0 = (AGG_PF - (AGG_LAG + 1) % AGG_PF) % AGG_PF;
static_assert(AGG_RING % AGG_PF == 0 && AGG_RING >= 2 * kMaxArm + 2, "ring geometry");
static_assert(AGG_PF <= AGG_LAG, "arm prefetch must stay inside the line");

// ---- prefix rings ----------------------------------------------------------------------------
// A ring holds the last AGG_RING prefixes of the thread's AGG_NC chains.  Two homes:
//   SmemRing<T>: shared memory, [AGG_RING][T threads] x 16 bytes (slot stride T*16 bytes);
//   TmemRing   : tensor memory.  A warp with warp%4 == q owns TMEM lanes 32q..32q+31, thread = lane,
//                slot s = 4 consecutive 32-bit columns 4s..4s+3 of that lane (tcgen05.st/ld 32x32b.x4).
//                Slot offsets must be warp-uniform there: every lane of the warp walks the SAME line.
// Offsets handed to st/ld are in ring units (SLOT per slot, SPAN per ring).
template <int THREADS>
struct SmemRing {
    static constexpr int SLOT = THREADS * 8 * AGG_NC;
    static constexpr int SPAN = AGG_RING * SLOT;
    struct Raw { double a, b; };
    uint32_t base;
    __device__ __forceinline__ void st(uint32_t off, double a, double b) const
    {
        if (AGG_NC == 2) asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(base + off), "d"(a), "d"(b) : "memory");
        else asm volatile("st.shared.f64 [%0], %1;" ::"r"(base + off), "d"(a) : "memory");
    }
    __device__ __forceinline__ void ld(uint32_t off, Raw& r) const
    {
        if (AGG_NC == 2) asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(r.a), "=d"(r.b) : "r"(base + off) : "memory");
        else { asm volatile("ld.shared.f64 %0, [%1];" : "=d"(r.a) : "r"(base + off) : "memory"); r.b = 0.0; }
    }
    __device__ __forceinline__ void st_fence() const {}
    __device__ __forceinline__ void ld_fence() const {}
    __device__ __forceinline__ void unpack(Raw& r, double& a, double& b) const { a = r.a; b = r.b; }
};

struct TmemRing {
    static_assert(AGG_NC == 2, "a TMEM slot is one fp64 pair");
    static constexpr int SLOT = 4;  // columns
    static constexpr int SPAN = AGG_RING * SLOT;
    struct Raw { uint32_t w[4]; };
    uint32_t base;  // (lane quarter << 16) | first column
    __device__ __forceinline__ void st(uint32_t off, double a, double b) const
    {
        uint32_t a0, a1, b0, b1;
        asm("mov.b64 {%0, %1}, %2;" : "=r"(a0), "=r"(a1) : "d"(a));
        asm("mov.b64 {%0, %1}, %2;" : "=r"(b0), "=r"(b1) : "d"(b));
        asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(base + off), "r"(a0), "r"(a1),
                     "r"(b0), "r"(b1)
                     : "memory");
    }
    __device__ __forceinline__ void ld(uint32_t off, Raw& r) const
    {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3])
                     : "r"(base + off)
                     : "memory");
    }
    // tcgen05.st / .ld are asynchronous: a store must be waited for before the slot is read back, and
    // the destination registers of a load must not be touched before wait::ld.
    __device__ __forceinline__ void st_fence() const { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
    __device__ __forceinline__ void ld_fence() const { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
    __device__ __forceinline__ void unpack(Raw& r, double& a, double& b) const
    {
        // volatile pins keep every use of the loaded words behind the (volatile) wait::ld
#pragma unroll
        for (int i = 0; i < 4; ++i) asm volatile("" : "+r"(r.w[i]));
        asm("mov.b64 %0, {%1, %2};" : "=d"(a) : "r"(r.w[0]), "r"(r.w[1]));
        asm("mov.b64 %0, {%1, %2};" : "=d"(b) : "r"(r.w[2]), "r"(r.w[3]));
    }
};

// Streaming accesses of the cost volume: every cell is touched exactly once per pass, so the
// lines must not occupy the (small, shared-memory-carved) L1.
__device__ __forceinline__ float2 ld_stream(const char* p)
{
    float2 v;
    if (AGG_NC == 2) asm volatile("ld.global.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    else { asm volatile("ld.global.L1::no_allocate.f32 %0, [%1];" : "=f"(v.x) : "l"(p)); v.y = 0.f; }
    return v;
}
// Single-pass walk: batches the L2 prefetch runs ahead of the register loads (0 = off).  All global loads of a warp share
// one or two scoreboard slots, so register buffers cannot cover the DRAM latency (see k_agg_fused); a prefetch has no
// destination register.  Measured at C3 (ms, horizontal pass / normalising horizontal pass; 6 register buffers unless noted):
// off 1.35 / 1.45, 4: 1.11 / 1.31, 6: 1.10 / 1.32, 8: 1.11 / 1.27, 10: 1.13 / 1.25, 16: 1.15 / 1.25; 8 with 4 buffers 1.10 / 1.27
// (kept: smaller loop), 8 with 2 buffers 1.20 / 1.26.
#ifndef TSM_AGG_PD
#define TSM_AGG_PD 8
#endif
constexpr int AGG_PD = TSM_AGG_PD;
__device__ __forceinline__ void prefetch_l2_lt(const char* p, int pos, int len)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.lt.u32 p, %1, %2;\n\t@p prefetch.global.L2 [%0];\n\t}" ::"l"(p), "r"(pos), "r"(len));
}
__device__ __forceinline__ void st_stream(char* p, float a, float b)
{

    if (AGG_NC == 2) asm volatile("st.global.L1::no_allocate.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
    else asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(p), "f"(a) : "memory");
}

static_assert(AGG_PF % 4 == 0, "descriptors are fetched four at a time; batches run in sub-blocks of 4");

// One thread walks its AGG_NC chains along one line.  `cell` = first cell of the chains, `cstride` =
// BYTES between consecutive positions (32-bit: keeps the pointer stepping to one add-with-carry), `desc_line` = the line's step descriptors.
// Requires len >= AGG_LAG + 1 + AGG_U (host-checked) and AGG_U positions of over-read slack
// behind every line end (the volumes are allocated with it).
template <bool NORM, int NB, class Ring>
__device__ __forceinline__ void walk_line(const Ring ring, float* cell, const uint32_t cstride, const uint32_t* desc_line,
                                          const float* rcp_line, const int len)
{
    constexpr int SLOT = Ring::SLOT, SPAN = Ring::SPAN;
    typedef typename Ring::Raw Raw;
    const char* in_ptr = reinterpret_cast<const char*>(cell);
    char* out_ptr = reinterpret_cast<char*>(cell);
    const uint32_t* desc_ptr = desc_line;

    double P0 = 0.0, P1 = 0.0;
    ring.st(AGG_P0 * SLOT, 0.0, 0.0);  // P[0]

    // ---- fill: pushes t = 0 .. AGG_LAG-1 (P[1..33] -> slots AGG_P0+1 .. AGG_P0+33, no wrap),
    // values fetched in 3 batches of 11 ----
    uint32_t hs = AGG_P0 * SLOT;  // slot of the newest prefix
    static_assert(AGG_LAG == 33, "fill batches");
#pragma unroll 1
    for (int g = 0; g < 3; ++g) {
        float2 tmp[11];
#pragma unroll
        for (int u = 0; u < 11; ++u) {
            tmp[u] = ld_stream(in_ptr);
            in_ptr += cstride;
        }
#pragma unroll
        for (int u = 0; u < 11; ++u) {
            P0 += (double)tmp[u].x;
            P1 += (double)tmp[u].y;
            hs += SLOT;
            ring.st(hs, P0, P1);
        }
    }
    // newest = P[33] at slot AGG_P0 + 33; the next prefix P[34] goes to a multiple of AGG_PF.
    uint32_t nx = hs + SLOT;  // slot of the next prefix to be written, multiple of AGG_PF slots
    if (nx == SPAN) nx = 0;

    // ---- main: steps t = AGG_LAG .. len-1: push in[t] -> P[t+1], emit o = t - AGG_LAG ----
    // Software pipeline in BATCHES of AGG_PF steps over NB register buffers: all loads of a
    // batch are issued back to back, NBUF-1 batches ahead of their use.  Batching matters: a warp
    // has only six scoreboard slots and a slot completes when ALL loads charged to it have landed,
    // so independent loads must be grouped by the time they are needed, not interleaved one per step.
    float2 vin[NB][AGG_PF];
    uint32_t av[NB][AGG_PF];
    // Address of step u of a batch = base + u*cstride as ONE 32x32+64 multiply-add; done in asm because
    // the compiler otherwise re-associates the unrolled loop into base + k*cstride with all the multiples
    // hoisted as loop invariants (77 extra live registers in the vertical pass).
    auto step_addr = [&](const char* base, int u) {
        const char* a;
        asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(a) : "r"(cstride), "r"((uint32_t)u), "l"(base));
        return a;
    };
    int t_load = AGG_LAG;  // position of in_ptr on the line
    auto load_batch = [&](int buf) {
#pragma unroll
        for (int u = 0; u < AGG_PF; ++u) vin[buf][u] = ld_stream(step_addr(in_ptr, u));
        if (AGG_PD > 0) {
            // the DRAM latency is covered by an L2 prefetch (no destination register, no scoreboard slot) AGG_PD batches ahead
#pragma unroll
            for (int u = 0; u < AGG_PF; ++u) prefetch_l2_lt(step_addr(in_ptr, AGG_PD * AGG_PF + u), t_load + AGG_PD * AGG_PF + u, len);
            t_load += AGG_PF;
        }
        in_ptr = step_addr(in_ptr, AGG_PF);
        asm volatile("" : "+l"(in_ptr));
#pragma unroll
        for (int j = 0; j < AGG_PF / 4; ++j) {
            const uint4 q = *reinterpret_cast<const uint4*>(desc_ptr + 4 * j);
            av[buf][4 * j + 0] = q.x; av[buf][4 * j + 1] = q.y; av[buf][4 * j + 2] = q.z; av[buf][4 * j + 3] = q.w;
        }
        desc_ptr += AGG_PF;
    };
    // One batch = AGG_PF steps, processed in sub-blocks of AGG_SUB: (A) the prefix pushes and their
    // ring stores, (B) all ring loads, (C) the outputs.  A warp issues in order, so finishing step u
    // right after its own ring loads would expose the ring-load -> DADD -> F2F -> STG latency on
    // every step; grouped like this it is paid once per sub-block.
    // Reciprocals of the divisors (normalising passes): L1-resident broadcast loads, fetched one batch ahead.
    float yv[2][AGG_PF];
    const float* rcp_ptr = rcp_line;
    auto load_rcp = [&](int par) {
        if (NORM) {
#pragma unroll
            for (int j = 0; j < AGG_PF / 4; ++j) {
                const float4 q = *reinterpret_cast<const float4*>(rcp_ptr + 4 * j);
                yv[par][4 * j + 0] = q.x; yv[par][4 * j + 1] = q.y; yv[par][4 * j + 2] = q.z; yv[par][4 * j + 3] = q.w;
            }
            rcp_ptr += AGG_PF;
        }
    };
    auto run_batch = [&](int buf, int nsteps) {  // nsteps == AGG_PF in the steady state
        const int par = buf & 1;
        load_rcp(par ^ 1);  // next batch
#pragma unroll
        for (int u0 = 0; u0 < AGG_PF; u0 += AGG_SUB) {
            Raw hi[AGG_SUB], lo[AGG_SUB];
#pragma unroll
            for (int w = 0; w < AGG_SUB; ++w) {
                const int u = u0 + w;
                if (u < nsteps) {
                    P0 += (double)vin[buf][u].x;
                    P1 += (double)vin[buf][u].y;
                    ring.st(nx + u * SLOT, P0, P1);
                }
            }
            ring.st_fence();
#pragma unroll
            for (int w = 0; w < AGG_SUB; ++w) {
                const int u = u0 + w;
                if (u < nsteps) {
                    const uint32_t desc = av[buf][u];
                    const int a = desc & 0xff, b = (desc >> 8) & 0xff;
                    const uint32_t top = nx + u * SLOT;
                    // slot offsets lie in (-SPAN, SPAN): a negative one wraps to a huge unsigned value,
                    // so the unsigned minimum with the offset + SPAN is the wrapped offset
                    const uint32_t s1 = top - (uint32_t)(AGG_LAG - b) * SLOT;      // P[o + b + 1]
                    const uint32_t s0 = top - (uint32_t)(AGG_LAG + 1 + a) * SLOT;  // P[o - a]
                    ring.ld(min(s1, s1 + SPAN), hi[w]);
                    ring.ld(min(s0, s0 + SPAN), lo[w]);
                }
            }
            ring.ld_fence();
#pragma unroll
            for (int w = 0; w < AGG_SUB; ++w) {
                const int u = u0 + w;
                if (u < nsteps) {
                    double h0, h1, l0, l1;
                    ring.unpack(hi[w], h0, h1);
                    ring.unpack(lo[w], l0, l1);
                    float r0 = __double2float_rn(h0 - l0), r1 = __double2float_rn(h1 - l1);
                    if (NORM) {
                        // N < 2^23: (float)N without the conversion pipe
                        const float nf = __fsub_rn(__uint_as_float(0x4B000000u | (av[buf][u] >> 16)), 8388608.f);
                        r0 = div_exact_rn(r0, nf, yv[par][u]);
                        r1 = div_exact_rn(r1, nf, yv[par][u]);
                    }
                    st_stream(const_cast<char*>(step_addr(out_ptr, u)), r0, r1);
                }
            }
        }
        out_ptr = const_cast<char*>(step_addr(out_ptr, nsteps));
        asm volatile("" : "+l"(out_ptr));
    };
    const int nB = len - AGG_LAG;
    uint32_t newest = hs;  // slot of the newest prefix
    int done = 0;
    // NB register buffers: NBUF-1 batches are in flight while one is being processed.
    static_assert(NB % 2 == 0, "reciprocal double buffer follows the batch parity");
#pragma unroll
    for (int b = 0; b < NB - 1; ++b) load_batch(b);
    load_rcp(0);
    for (; done + NB * AGG_PF <= nB; done += NB * AGG_PF) {
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            load_batch((b + NB - 1) % NB);
            run_batch(b, AGG_PF);
            nx += AGG_PF * SLOT;
            if (nx == SPAN) nx = 0;
        }
    }
    // tail: fewer than NBUF*PF steps left; buffers 0 .. NBUF-2 already hold the next batches
    {
        int rem = nB - done;
        newest = (nx == 0 ? SPAN : nx) - SLOT;
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            if (rem > 0) {
                if (b == NB - 1) load_batch(b);  // the one buffer that was not prefetched
                const int n = rem < AGG_PF ? rem : AGG_PF;
                run_batch(b, n);
                newest = nx + (uint32_t)(n - 1) * SLOT;
                nx += AGG_PF * SLOT;
                if (nx == SPAN) nx = 0;
                rem -= n;
            }
        }
    }
    // ---- drain: o = len-LAG .. len-1.  Newest prefix stays P[len]; the virtual slot of P[o+34]
    // is (j+1) slots past it.  Only real slots (<= o+b+1 <= len) are read.
#pragma unroll 1
    for (int j = 0; j < AGG_LAG; ++j) {
        const int o = len - AGG_LAG + j;
        uint32_t top = newest + (uint32_t)(j + 1) * SLOT;
        top -= (top >= SPAN) ? SPAN : 0;
        const uint32_t desc = desc_line[o];
        const int a = desc & 0xff, b = (desc >> 8) & 0xff;
        int s1 = (int)top - (AGG_LAG - b) * SLOT;      // P[o + b + 1]
        int s0 = (int)top - (AGG_LAG + 1 + a) * SLOT;  // P[o - a]
        s1 += (s1 < 0) ? SPAN : 0;
        s0 += (s0 < 0) ? SPAN : 0;
        Raw hi, lo;
        ring.ld(s1, hi);
        ring.ld(s0, lo);
        ring.ld_fence();
        double h0, h1, l0, l1;
        ring.unpack(hi, h0, h1);
        ring.unpack(lo, l0, l1);
        float r0 = __double2float_rn(h0 - l0), r1 = __double2float_rn(h1 - l1);
        if (NORM) {
            const float nf = (float)(desc >> 16), y = rcp_line[o];
            r0 = div_exact_rn(r0, nf, y);
            r1 = div_exact_rn(r1, nf, y);
        }
        st_stream(out_ptr, r0, r1);
        out_ptr += cstride;
    }
}

// Which chains a thread owns: chain index -> (part, line, first disparity).
struct ChainSel {
    float* cell;
    uint32_t cstride;  // bytes
    const uint32_t* desc_line;
    const float* rcp_line;
};
template <bool VERT>
__device__ __forceinline__ ChainSel select_chain(const Dims& dm, const ViewPtrs& v, bool tail_part, long long chain, int npair)
{
    float* part = tail_part ? v.vol.tail : v.vol.main;
    const int pitch = tail_part ? dm.Rp : dm.Dm;
    const int line = (int)(chain / npair), d = AGG_NC * (int)(chain % npair);
    const size_t line_px = VERT ? (size_t)line : (size_t)line * dm.W;
    ChainSel c;
    c.cell = part + line_px * pitch + d;
    c.cstride = (VERT ? (uint32_t)dm.W * (uint32_t)pitch : (uint32_t)pitch) * 4u;
    c.desc_line = VERT ? v.desc_v + (size_t)line * dm.Hd() : v.desc_h + (size_t)line * dm.Wd();
    c.rcp_line = VERT ? v.rcp_v + (size_t)line * dm.Hd() : v.rcp_h + (size_t)line * dm.Wd();
    return c;
}

// ---- kernel 1: shared-memory rings only (any Dm) ----
// 64-thread CTAs, 73.7 KB of ring each, 3 CTAs = 6 warps per SM.
template <bool VERT, bool NORM>
__global__ void __launch_bounds__(AGG_BLOCK)
k_agg_walk(Dims dm, ViewPtrs v0, ViewPtrs v1, int nb_main)
{
    extern __shared__ __align__(16) unsigned char ring_raw[];  // [AGG_RING][AGG_BLOCK] x double2
    const ViewPtrs& v = blockIdx.y ? v1 : v0;
    const int nlines = VERT ? dm.W : dm.H, len = VERT ? dm.H : dm.W;
    // chains of the 128-byte aligned main part first, then (blocks >= nb_main) those of the tail part
    const bool tail_part = (int)blockIdx.x >= nb_main;
    const int npair = tail_part ? (dm.tail() + AGG_NC - 1) / AGG_NC : dm.Dm / AGG_NC;
    const long long chain = (long long)((int)blockIdx.x - (tail_part ? nb_main : 0)) * AGG_BLOCK + threadIdx.x;
    if (chain >= (long long)nlines * npair) return;
    const ChainSel c = select_chain<VERT>(dm, v, tail_part, chain, npair);
    SmemRing<AGG_BLOCK> ring;
    ring.base = (uint32_t)__cvta_generic_to_shared(ring_raw) + threadIdx.x * (8 * AGG_NC);
    walk_line<NORM, AGG_NBUF_STATIC>(ring, c.cell, c.cstride, c.desc_line, c.rcp_line, len);
}

// ---- kernel 2: persistent, shared-memory AND tensor-memory rings (Dm % 64 == 0) ----
// The walk is latency-bound per warp and its throughput grows with the number of resident warps,
// which the 1152-byte ring per thread caps at 6 per SM in shared memory.  The SM's 256 KB of
// tensor memory is idle in this pipeline, so it takes up to four more rings-worth of warps: one CTA
// per SM, warps 0..tm-1 keep their rings in TMEM (one lane quarter each, 288 of 512 columns),
// the others in 216 KB of shared memory.  TMEM slot addresses are warp-uniform, which holds when a
// warp's 32 threads walk the same line: a work item is (view, line, 32 adjacent chain pairs = 64
// disparities), hence Dm % 64 == 0.
// Work items are handed out through a global counter, one warp at a time (no CTA-wide barrier
// in the loop): a line walk is long (hundreds of microseconds), a static split leaves 10 % of the
// SM-time idle in the last wave and would couple the faster and slower ring homes.
// Tail part (Dn - Dm < 32 disparities per pixel), vertical passes: items of 32 columns x one pair, lanes
// over columns; slot addresses are then per-lane, so only shared-memory warps take them -- first, so that
// they overlap with the main items.  Horizontal passes of the tail part: k_agg_tail_h.
constexpr int AGH_TM_MAX = 4, AGH_SM_WARPS = 6;
constexpr int AGH_SM_THREADS = 32 * AGH_SM_WARPS;
constexpr int AGH_SMEM = SmemRing<AGH_SM_THREADS>::SPAN;
constexpr int AGH_MAX_BLOCK = 32 * (AGH_TM_MAX + AGH_SM_WARPS);
static_assert(TmemRing::SPAN <= 512, "ring fits the TMEM columns");

__device__ __forceinline__ unsigned next_item(unsigned* ctr, int lane)
{
    unsigned v = 0;
    if (lane == 0) v = atomicAdd(ctr, 1u);
    return __shfl_sync(0xffffffffu, v, 0);
}

template <bool VERT, bool NORM>
__global__ void __launch_bounds__(AGH_MAX_BLOCK, 1)
k_agg_persist(Dims dm, ViewPtrs v0, ViewPtrs v1, int tm_warps, unsigned* ctr)
{
    extern __shared__ __align__(16) unsigned char ring_raw[];  // [AGG_RING][AGH_SM_THREADS] x double2
    __shared__ uint32_t tm_base;
    const int nlines = VERT ? dm.W : dm.H, len = VERT ? dm.H : dm.W;
    // warp index through a lane-0 broadcast: tells the compiler it is warp-uniform (TMEM addresses live in uniform registers)
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    if (tm_warps > 0) {
        if (warp == 0) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(
                             (uint32_t)__cvta_generic_to_shared(&tm_base))
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    const unsigned ngroups = dm.Dm / 64, per_view = (unsigned)nlines * ngroups, n_main = 2 * per_view;
    const size_t line_step = VERT ? (size_t)dm.Dm : (size_t)dm.W * dm.Dm;  // floats between lines, main part
    const uint32_t cstride_main = (VERT ? (uint32_t)dm.W * (uint32_t)dm.Dm : (uint32_t)dm.Dm) * 4u;
    const int desc_pitch = VERT ? dm.Hd() : dm.Wd();
    if (warp < tm_warps) {
        TmemRing ring;
        ring.base = tm_base + ((uint32_t)(warp * 32) << 16);
        unsigned item = next_item(ctr, lane);
        while (item < n_main) {
            const unsigned nxt = next_item(ctr, lane);  // its latency hides behind the walk
            const ViewPtrs& v = item >= per_view ? v1 : v0;
            const unsigned r = item >= per_view ? item - per_view : item;
            const unsigned line = r / ngroups, grp = r - line * ngroups;
            float* cell = v.vol.main + line * line_step + 64 * grp + 2 * lane;
            const uint32_t* desc_line = (VERT ? v.desc_v : v.desc_h) + (size_t)line * desc_pitch;
            const float* rcp_line = (VERT ? v.rcp_v : v.rcp_h) + (size_t)line * desc_pitch;
            walk_line<NORM, AGG_NBUF_PERSIST>(ring, cell, cstride_main, desc_line, rcp_line, len);
            item = nxt;
        }
    } else {
        SmemRing<AGH_SM_THREADS> ring;
        ring.base = (uint32_t)__cvta_generic_to_shared(ring_raw) + (threadIdx.x - 32 * tm_warps) * (8 * AGG_NC);
        // tail part first
        const int tpairs = (dm.tail() + AGG_NC - 1) / AGG_NC;
        const unsigned lgroups = (unsigned)(nlines + 31) / 32, tail_per_view = lgroups * tpairs;
        const unsigned n_tail = VERT ? 2 * tail_per_view : 0;  // horizontal: k_agg_tail_h
        if (n_tail) {
            const size_t tline_step = VERT ? (size_t)dm.Rp : (size_t)dm.W * dm.Rp;
            const uint32_t cstride_tail = (VERT ? (uint32_t)dm.W * (uint32_t)dm.Rp : (uint32_t)dm.Rp) * 4u;
            unsigned item = next_item(ctr + 1, lane);
            while (item < n_tail) {
                const unsigned nxt = next_item(ctr + 1, lane);
                const ViewPtrs& v = item >= tail_per_view ? v1 : v0;
                const unsigned r = item >= tail_per_view ? item - tail_per_view : item;
                const unsigned lg = r / tpairs, pr = r - lg * tpairs;
                const unsigned line = 32 * lg + lane;
                if (line < (unsigned)nlines) {
                    float* cell = v.vol.tail + line * tline_step + AGG_NC * pr;
                    const uint32_t* desc_line = (VERT ? v.desc_v : v.desc_h) + (size_t)line * desc_pitch;
                    const float* rcp_line = (VERT ? v.rcp_v : v.rcp_h) + (size_t)line * desc_pitch;
                    walk_line<NORM, AGG_NBUF_PERSIST>(ring, cell, cstride_tail, desc_line, rcp_line, len);
                }
                item = nxt;
            }
        }
        unsigned item = next_item(ctr, lane);
        while (item < n_main) {
            const unsigned nxt = next_item(ctr, lane);
            const ViewPtrs& v = item >= per_view ? v1 : v0;
            const unsigned r = item >= per_view ? item - per_view : item;
            const unsigned line = r / ngroups, grp = r - line * ngroups;
            float* cell = v.vol.main + line * line_step + 64 * grp + 2 * lane;
            const uint32_t* desc_line = (VERT ? v.desc_v : v.desc_h) + (size_t)line * desc_pitch;
            const float* rcp_line = (VERT ? v.rcp_v : v.rcp_h) + (size_t)line * desc_pitch;
            walk_line<NORM, AGG_NBUF_PERSIST>(ring, cell, cstride_main, desc_line, rcp_line, len);
            item = nxt;
        }
    }
    if (tm_warps > 0) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm_base) : "memory");
    }
}


// ---- kernel 4: TWO consecutive same-direction passes in one walk (V.norm -> V, H.norm -> H) ----
// The pass order of costAggregate is H, V.n | V, H.n | H, V.n | V, H.n (ADCensus.cpp:774-783; .n = the pass that ends
// an iteration and divides by the window size): three adjacent pairs walk the same lines with the same arms, so one
// walk can apply both -- the volume makes 5 round trips through HBM instead of 8.  Per chain that takes two prefix
// rings: pass A (normalising) emits c(o1) = (PA[o1+b+1] - PA[o1-a]) / N(o1) with a lag of 33 behind the input
// position t; pass B pushes c(o1) into its own prefix PB and emits out(o2) = PB[o2+b+1] - PB[o2-a] another 33
// (+ one batch of 4: B works on the previous batch's c so that the two ring round trips overlap) behind.
//
// Ring storage decides how many chains an SM can keep in flight, and the walk needs warps in flight to cover the HBM
// latency.  Two fp64 rings per chain leave room for 5 warps.  The rings here hold 48-bit WRAP-AROUND FIXED-POINT
// prefixes instead: only differences of prefixes at most 67 positions apart are ever used, and those are < 2^14
// (normalising pass: a window of <= 67 first-pass sums <= 134) resp. < 2^8 (plain pass: <= 67 averaged costs <= 2),
// so 48 bits mod 2^48 carry them exactly at a resolution of 2^-34 resp. 2^-40 -- the resolution an fp64 prefix has at
// the far end of a 1920-pixel line, and uniform along the line.  A slot is 12 bytes per chain pair (two low words +
// the two 16-bit high halves in one word): pass A's ring is 216 of the 256 tensor-memory columns a warp can call its
// own when eight warps share the four lane quarters, pass B's ring 27 KB of shared memory per warp -- 8 warps per SM
// with both rings, all running the same code.  scripts/emulate_agg_fixed.py checks the arithmetic against the oracle.
//
// The walk is branch-free over garbage: it starts 3 positions in front of the line (so that pass A's descriptor
// groups are 16-byte aligned), runs 70 positions past its end, loads outside [0, len) are predicated off, stores
// outside [0, len) likewise, and whatever the prefixes collect before position 0 is a common offset that cancels in
// every difference (the arithmetic is modular, so it cancels exactly).  Descriptors read before / behind a line are
// the neighbouring lines' or zero padding: any of them keeps the ring offsets inside the ring.
#ifndef TSM_AGF_NBUF
#define TSM_AGF_NBUF 3
#endif
#ifndef TSM_AGF_PD
#define TSM_AGF_PD 10
#endif
constexpr int AGF_WARPS = 8, AGF_THREADS = 32 * AGF_WARPS;
constexpr int AGF_NB = TSM_AGF_NBUF;             // register buffers of AGG_PF input positions each
constexpr int AGF_PD = TSM_AGF_PD;               // batches the L2 prefetch runs ahead of the register loads (0 = off)
constexpr int AGF_SPAN3 = 3 * AGG_RING;          // ring positions are kept in units of 3 = TMEM columns per slot
constexpr int AGF_SMEM = AGG_RING * 3 * 1024;    // pass B rings: [slot][256 x 8 B low words | 256 x 4 B high halves]
constexpr int AGF_LEAD = 3;                      // positions walked in front of the line
constexpr int AGF_LAG_A = AGG_LAG;               // o1 = t - 33
constexpr int AGF_LAG_B = 2 * AGG_LAG + AGG_PF;  // o2 = t - 70
constexpr int AGF_SA = 34, AGF_SB = 40;          // fixed-point scales of the two passes
static_assert(AGG_PF == 4 && AGG_NC == 2 && AGG_RING % 4 == 0, "fused walk geometry");
static_assert((AGF_LEAD + AGF_LAG_A) % 4 == 0, "pass A descriptor groups are aligned");
static_assert(AGF_SPAN3 + 1 <= 256, "a warp's tensor-memory share");

// low 48 bits = round(v * 2^S) mod 2^48; the bits above are junk that the rings drop
template <int S>
__device__ __forceinline__ uint64_t fix48(float v)
{
    constexpr double magic = 1.5 * (double)(1ull << (52 - S));
    // float -> double by re-biasing the exponent with two integer instructions instead of a conversion (F2F runs on the
    // quarter-rate XU pipe and every one of them occupies a scoreboard slot the loads need).  Exact for positive normal
    // numbers; +0 and denormals (< 2^-126) become a value below 2^-126, which the addition of `magic` rounds away just as
    // it would round the true value; anything else only occurs in discarded positions.
    const uint32_t b = __float_as_uint(v);
    const double w = __hiloint2double((int)((b >> 3) + 0x38000000u), (int)(b << 29));
    return (uint64_t)__double_as_longlong(__dadd_rn(w, magic));
}
// (hi - lo) mod 2^48 of two ring entries, times 2^-S, rounded once to fp32 -- for both chains of the pair
// `exp_word` = (1023 + 52 - S) << 20, handed in from a per-thread register: as a literal the compiler parks it in a uniform
// register and copies it into a vector register in front of every PRMT (15 extra instructions per batch).
template <int S>
__device__ __forceinline__ void diff48(const uint32_t* h, const uint32_t* l, float& r0, float& r1, uint32_t exp_word)
{
    const uint32_t EXP = exp_word;
    constexpr double base = (double)(1ull << (52 - S));
    uint32_t d0lo, d0hi, d1lo, d1hi;
    asm("sub.cc.u32 %0, %2, %3;\n\tsubc.u32 %1, %4, %5;" : "=r"(d0lo), "=r"(d0hi) : "r"(h[0]), "r"(l[0]), "r"(h[2]), "r"(l[2]));
    asm("sub.cc.u32 %0, %2, %3;\n\tsubc.u32 %1, %4, %5;"
        : "=r"(d1lo), "=r"(d1hi)
        : "r"(h[1]), "r"(l[1]), "r"(h[2] >> 16), "r"(l[2] >> 16));
    d0hi = __byte_perm(d0hi, EXP, 0x7610);  // (d & 0xffff) | EXP in one instruction
    d1hi = __byte_perm(d1hi, EXP, 0x7610);
    r0 = __double2float_rn(__dsub_rn(__hiloint2double((int)d0hi, (int)d0lo), base));
    r1 = __double2float_rn(__dsub_rn(__hiloint2double((int)d1hi, (int)d1lo), base));
}

__device__ __forceinline__ void ld_stream_if(float2& v, const char* p, int pos, int len)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.lt.u32 p, %3, %4;\n\t@p ld.global.L1::no_allocate.v2.f32 {%0, %1}, [%2];\n\t}"
                 : "+f"(v.x), "+f"(v.y)
                 : "l"(p), "r"(pos), "r"(len));
}
__device__ __forceinline__ void prefetch_l2_if(const char* p, int pos, int len)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.lt.u32 p, %1, %2;\n\t@p prefetch.global.L2 [%0];\n\t}" ::"l"(p), "r"(pos), "r"(len));
}
__device__ __forceinline__ void st_stream_if(char* p, float a, float b, int pos, int len)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.lt.u32 p, %3, %4;\n\t@p st.global.L1::no_allocate.v2.f32 [%0], {%1, %2};\n\t}" ::"l"(p),
                 "f"(a), "f"(b), "r"(pos), "r"(len)
                 : "memory");
}

// One thread walks its two chains along one line through both passes.  tm = tensor-memory address of the warp's
// pass-A ring (lane quarter << 16 | first column), sl / sh = shared-memory addresses of the thread's pass-B ring
// (low-word pairs, high halves; a slot is 3 KB further).
__device__ __forceinline__ void walk_fused(const uint32_t tm, const uint32_t sl, const uint32_t sh, float* cell, const uint32_t cstride,
                                           const uint32_t* fdesc_line, const float* rcp_line, const int len)
{
    auto step_addr = [&](const char* base, int u) {
        const char* a;
        asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(a) : "r"(cstride), "r"((uint32_t)u), "l"(base));
        return a;
    };
    // position of the first step: t = -AGF_LEAD
    const char* in_ptr = reinterpret_cast<const char*>(cell) - (size_t)AGF_LEAD * cstride;
    char* out_ptr = reinterpret_cast<char*>(cell) - (size_t)(AGF_LEAD + AGF_LAG_B) * cstride;
    const uint32_t* da_ptr = fdesc_line - (AGF_LEAD + AGF_LAG_A);   // pass A: positions t - 33, 16-byte aligned groups
    const float* ry_ptr = rcp_line - (AGF_LEAD + AGF_LAG_A);
    const uint32_t* db_ptr = fdesc_line - (AGF_LEAD + AGF_LAG_B - 1);  // pass B: the group holding positions t - 69 .. t - 66 of u = 0;
                                                                       // its first three words + the last word of the group before
    int t_load = -AGF_LEAD;            // position of the next batch to be loaded
    int o_out = -(AGF_LEAD + AGF_LAG_B);  // output position of pass B for u = 0 of the batch being processed

    // Inputs AND side data (descriptors of both passes, reciprocals) of a batch are loaded together, AGF_NB - 1 batches
    // ahead of their use: a warp has six scoreboard slots and a slot completes when ALL loads charged to it have landed,
    // so loads are grouped by the time they are needed (side data fetched one batch ahead shared slots with younger
    // volume loads and stalled on them: 25 % of the samples in the first ncu capture).
    float2 vin[AGF_NB][AGG_PF];
    uint32_t fa[AGF_NB][AGG_PF], fb[AGF_NB][AGG_PF];
    float yv[AGF_NB][AGG_PF];
    uint32_t fb_carry = 0;
#pragma unroll
    for (int b = 0; b < AGF_NB; ++b)
#pragma unroll
        for (int u = 0; u < AGG_PF; ++u) vin[b][u] = make_float2(0.f, 0.f);
    // All global loads of a warp end up on one or two of its six scoreboard slots (ptxas), so waiting for the oldest
    // batch also waits for the youngest: register prefetching deeper than one batch buys nothing.  The DRAM latency is
    // covered by an L2 prefetch (no destination register, no scoreboard) AGF_PD batches ahead instead; the register
    // loads then hit L2 one or two batches ahead of their use.
    const char* pf_ptr = in_ptr + (size_t)(AGF_PD * AGG_PF) * cstride;
    auto load_batch = [&](int buf) {
        if (AGF_PD > 0) {
#pragma unroll
            for (int u = 0; u < AGG_PF; ++u) prefetch_l2_if(step_addr(pf_ptr, u), t_load + AGF_PD * AGG_PF + u, len);
            pf_ptr = step_addr(pf_ptr, AGG_PF);
            asm volatile("" : "+l"(pf_ptr));
        }
#pragma unroll
        for (int u = 0; u < AGG_PF; ++u) ld_stream_if(vin[buf][u], step_addr(in_ptr, u), t_load + u, len);
        in_ptr = step_addr(in_ptr, AGG_PF);
        asm volatile("" : "+l"(in_ptr));
        t_load += AGG_PF;
        const uint4 qa = *reinterpret_cast<const uint4*>(da_ptr);
        const float4 qy = *reinterpret_cast<const float4*>(ry_ptr);
        const uint4 qb = *reinterpret_cast<const uint4*>(db_ptr);
        fa[buf][0] = qa.x; fa[buf][1] = qa.y; fa[buf][2] = qa.z; fa[buf][3] = qa.w;
        yv[buf][0] = qy.x; yv[buf][1] = qy.y; yv[buf][2] = qy.z; yv[buf][3] = qy.w;
        fb[buf][0] = fb_carry; fb[buf][1] = qb.x; fb[buf][2] = qb.y; fb[buf][3] = qb.z;
        fb_carry = qb.w;
        da_ptr += AGG_PF; ry_ptr += AGG_PF; db_ptr += AGG_PF;
    };

    uint64_t PA0 = 0, PA1 = 0, PB0 = 0, PB1 = 0;
    uint32_t topA = 0, topB = AGF_SPAN3 - 3 * AGG_PF;  // ring position (x3) of the newest batch of each pass; B trails A by one batch
    // (len >> 31 is 0, but neither nvcc nor ptxas can fold it: the words stay in per-thread registers)
    const uint32_t exp_a = ((uint32_t)(1023 + 52 - AGF_SA) << 20) | ((uint32_t)len >> 31);
    const uint32_t exp_b = ((uint32_t)(1023 + 52 - AGF_SB) << 20) | ((uint32_t)len >> 31);

    // pass A pushes of one batch: prefixes of the inputs in `buf`, packed for tensor memory (no memory traffic here)
    auto push_a = [&](int buf, uint32_t (&wa)[3 * AGG_PF]) {
#pragma unroll
        for (int u = 0; u < AGG_PF; ++u) {
            PA0 += fix48<AGF_SA>(vin[buf][u].x);
            PA1 += fix48<AGF_SA>(vin[buf][u].y);
            wa[3 * u + 0] = (uint32_t)PA0;
            wa[3 * u + 1] = (uint32_t)PA1;
            wa[3 * u + 2] = __byte_perm((uint32_t)(PA0 >> 32), (uint32_t)(PA1 >> 32), 0x5410);
        }
    };
    auto store_a = [&](uint32_t top, const uint32_t (&wa)[3 * AGG_PF]) {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(tm + top), "r"(wa[0]),
                     "r"(wa[1]), "r"(wa[2]), "r"(wa[3]), "r"(wa[4]), "r"(wa[5]), "r"(wa[6]), "r"(wa[7])
                     : "memory");
        asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(tm + top + 8), "r"(wa[8]), "r"(wa[9]),
                     "r"(wa[10]), "r"(wa[11])
                     : "memory");
    };

    // One loop body = batch k of pass A and batch k - 1 of pass B, software-pipelined so that no ring round trip is
    // exposed: the pushes of batch k are already in tensor memory (previous body); the ring loads of both passes are
    // issued first, the prefixes of batch k + 1 are computed while they fly and stored once they have landed (the
    // stores reuse the oldest slots, which this batch still reads), the outputs follow, and pass A's outputs go into
    // pass B's ring at the end, to be read by the next body.
    auto run_batch = [&](int b) {
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        // ---- ring loads of both passes ----
        uint32_t ha[AGG_PF][4], la[AGG_PF][4], hb[AGG_PF][3], lb[AGG_PF][3];
#pragma unroll
        for (int u = 0; u < AGG_PF; ++u) {
            const uint32_t f = fa[b][u], top = topA + 3 * u;
            const uint32_t s1 = top - (f & 0xffu), s0 = top - __byte_perm(f, 0, 0x4441);  // P[o + b + 1], P[o - a]
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(ha[u][0]), "=r"(ha[u][1]), "=r"(ha[u][2]), "=r"(ha[u][3])
                         : "r"(tm + min(s1, s1 + AGF_SPAN3))
                         : "memory");
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(la[u][0]), "=r"(la[u][1]), "=r"(la[u][2]), "=r"(la[u][3])
                         : "r"(tm + min(s0, s0 + AGF_SPAN3))
                         : "memory");
        }
#pragma unroll
        for (int u = 0; u < AGG_PF; ++u) {
            const uint32_t f = fb[b][u], top = topB + 3 * u;
            uint32_t s1 = top - (f & 0xffu), s0 = top - __byte_perm(f, 0, 0x4441);
            s1 = min(s1, s1 + AGF_SPAN3) << 10;
            s0 = min(s0, s0 + AGF_SPAN3) << 10;
            asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(hb[u][0]), "=r"(hb[u][1]) : "r"(sl + s1) : "memory");
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(hb[u][2]) : "r"(sh + s1) : "memory");
            asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(lb[u][0]), "=r"(lb[u][1]) : "r"(sl + s0) : "memory");
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(lb[u][2]) : "r"(sh + s0) : "memory");
        }
        // ---- prefixes of the next batch of pass A while the ring loads are in flight ----
        uint32_t wa[3 * AGG_PF];
        push_a((b + 1) % AGF_NB, wa);
        uint32_t nextA = topA + 3 * AGG_PF;
        if (nextA == AGF_SPAN3) nextA = 0;
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int u = 0; u < AGG_PF; ++u)
#pragma unroll
            for (int i = 0; i < 3; ++i) {  // keep every use of the loaded words behind the (volatile) wait::ld
                asm volatile("" : "+r"(ha[u][i]));
                asm volatile("" : "+r"(la[u][i]));
            }
        store_a(nextA, wa);
        // ---- pass B outputs ----
#pragma unroll
        for (int u = 0; u < AGG_PF; ++u) {
            float r0, r1;
            diff48<AGF_SB>(hb[u], lb[u], r0, r1, exp_b);
            st_stream_if(const_cast<char*>(step_addr(out_ptr, u)), r0, r1, o_out + u, len);
        }
        // ---- pass A outputs c = window sum / N, pushed into pass B's ring for the next body ----
        const uint32_t slB = sl + (topA << 10), shB = sh + (topA << 10);  // pass B's ring position of this batch = pass A's label
#pragma unroll
        for (int u = 0; u < AGG_PF; ++u) {
            float r0, r1;
            diff48<AGF_SA>(ha[u], la[u], r0, r1, exp_a);
            const float nf = __fsub_rn(__uint_as_float(0x4B000000u | (fa[b][u] >> 16)), 8388608.f);  // (float)N, N < 2^16
            PB0 += fix48<AGF_SB>(div_exact_rn(r0, nf, yv[b][u]));
            PB1 += fix48<AGF_SB>(div_exact_rn(r1, nf, yv[b][u]));
            asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(slB + u * 3072), "r"((uint32_t)PB0), "r"((uint32_t)PB1) : "memory");
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(shB + u * 3072),
                         "r"(__byte_perm((uint32_t)(PB0 >> 32), (uint32_t)(PB1 >> 32), 0x5410))
                         : "memory");
        }
        out_ptr = const_cast<char*>(step_addr(out_ptr, AGG_PF));
        asm volatile("" : "+l"(out_ptr));
        o_out += AGG_PF;
        topB = topA;
        topA = nextA;
        load_batch(b);  // this buffer's inputs were consumed by the previous body, its side data by this one
    };

    // batches: positions -3 .. len + 69 (the last output of pass B is position len - 1 = t - 70)
    const int nbatch = (len + AGF_LAG_B - 1 + AGF_LEAD) / AGG_PF + 1;
    if (AGF_PD > 0) {
#pragma unroll 1
        for (int j = 0; j < AGF_PD * AGG_PF; ++j) prefetch_l2_if(step_addr(in_ptr, j), j - AGF_LEAD, len);
    }
#pragma unroll
    for (int b = 0; b < AGF_NB; ++b) load_batch(b);
    {
        uint32_t wa[3 * AGG_PF];
        push_a(0, wa);
        store_a(topA, wa);
    }
    int done = 0;
    for (; done + AGF_NB <= nbatch; done += AGF_NB) {
#pragma unroll
        for (int b = 0; b < AGF_NB; ++b) run_batch(b);
    }
    {
        int rem = nbatch - done;
#pragma unroll
        for (int b = 0; b < AGF_NB; ++b) {
            if (rem > 0) {
                run_batch(b);
                --rem;
            }
        }
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

template <bool VERT>
__global__ void __launch_bounds__(AGF_THREADS, 1)
k_agg_fused(Dims dm, ViewPtrs v0, ViewPtrs v1, unsigned* ctr)
{
    extern __shared__ __align__(16) unsigned char ring_raw[];
    __shared__ uint32_t tm_base;
    const int nlines = VERT ? dm.W : dm.H, len = VERT ? dm.H : dm.W;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(
                         (uint32_t)__cvta_generic_to_shared(&tm_base))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // eight warps share the four lane quarters of tensor memory: warp w owns lanes 32 (w & 3) .. + 31, columns 256 (w >> 2) .. + 255
    const uint32_t tm = tm_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(warp >> 2) * 256u;
    const uint32_t smem = (uint32_t)__cvta_generic_to_shared(ring_raw);
    const uint32_t sl = smem + threadIdx.x * 8, sh = smem + 2048 + threadIdx.x * 4;
    const unsigned ngroups = dm.Dm / 64, per_view = (unsigned)nlines * ngroups, n_items = 2 * per_view;
    const size_t line_step = VERT ? (size_t)dm.Dm : (size_t)dm.W * dm.Dm;  // floats between lines
    const uint32_t cstride = (VERT ? (uint32_t)dm.W * (uint32_t)dm.Dm : (uint32_t)dm.Dm) * 4u;
    const int desc_pitch = VERT ? dm.Hd() : dm.Wd();
    unsigned item = next_item(ctr, lane);
    while (item < n_items) {
        const unsigned nxt = next_item(ctr, lane);  // its latency hides behind the walk
        const ViewPtrs& v = item >= per_view ? v1 : v0;
        const unsigned r = item >= per_view ? item - per_view : item;
        const unsigned line = r / ngroups, grp = r - line * ngroups;
        float* cell = v.vol.main + line * line_step + 64 * grp + 2 * lane;
        const uint32_t* fdesc_line = (VERT ? v.fdesc_v : v.fdesc_h) + (size_t)line * desc_pitch;
        const float* rcp_line = (VERT ? v.rcp_v : v.rcp_h) + (size_t)line * desc_pitch;
        walk_fused(tm, sl, sh, cell, cstride, fdesc_line, rcp_line, len);
        item = nxt;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm_base) : "memory");
}

// ---- kernel 3: horizontal pass over the tail part ----
// In the tail part [H][W][Rp] a horizontal chain steps Rp floats at a time and neighbouring chains of a
// warp would be whole image rows apart (one 32-byte sector per lane and step: measured +0.6 ms per pass
// for 1/96 of the data).  A row of the tail is small, so here one CTA takes (row, view, disparity pair),
// stages the row in shared memory with coalesced loads, turns it into fp64 prefix sums with a block
// scan, and every output is the same  P[x+b+1] - P[x-a]  as in the walk.
constexpr int AGT_BLOCK = 256;

// In-place inclusive scan of A[1 .. n] (fp64 pairs) by one CTA: warp w owns the contiguous range [w*R, (w+1)*R) and walks it
// 32 elements at a time (consecutive lanes = consecutive 16-byte slots: conflict-free) with a running carry; the warp
// totals are exchanged through shared memory once and added in a second sweep.  Ends with a barrier.
__device__ __forceinline__ void tail_scan(double2* A, double2* warp_tot, int n)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int R = (((n + AGT_BLOCK / 32 - 1) / (AGT_BLOCK / 32)) + 31) & ~31;
    const int xb = warp * R, xe = min(n, xb + R);
    double c0 = 0.0, c1 = 0.0;
    for (int x = xb + lane; x - lane < xe; x += 32) {
        double i0 = 0.0, i1 = 0.0;
        if (x < xe) { const double2 p = A[x + 1]; i0 = p.x; i1 = p.y; }
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double t0 = __shfl_up_sync(0xffffffffu, i0, o), t1 = __shfl_up_sync(0xffffffffu, i1, o);
            if (lane >= o) { i0 += t0; i1 += t1; }
        }
        i0 += c0; i1 += c1;
        if (x < xe) A[x + 1] = make_double2(i0, i1);
        c0 = __shfl_sync(0xffffffffu, i0, 31);
        c1 = __shfl_sync(0xffffffffu, i1, 31);
    }
    __syncthreads();  // (also protects warp_tot against the previous scan's readers)
    if (lane == 0) warp_tot[warp] = make_double2(c0, c1);
    __syncthreads();
    double o0 = 0.0, o1 = 0.0;
    for (int w = 0; w < warp; ++w) { o0 += warp_tot[w].x; o1 += warp_tot[w].y; }
    if (warp > 0) {
        for (int x = xb + lane; x < xe; x += 32) {
            double2 p = A[x + 1];
            p.x += o0; p.y += o1;
            A[x + 1] = p;
        }
    }
    __syncthreads();
}

// MODE 0: one plain pass, 1: one normalising pass, 2: a normalising pass followed by a plain pass (the tail part's share
// of a fused launch).  One CTA per (line, view, disparity pair); the line is staged in shared memory.
template <bool VERT, int MODE>
__global__ void __launch_bounds__(AGT_BLOCK)
k_agg_tail(Dims dm, ViewPtrs v0, ViewPtrs v1)
{
    extern __shared__ __align__(16) unsigned char tail_raw[];
    const int len = VERT ? dm.H : dm.W;
    double2* P = reinterpret_cast<double2*>(tail_raw);  // [len + 1]
    double2* Q = P + (len + 1);                         // [len + 1], MODE 2 only
    __shared__ double2 warp_tot[AGT_BLOCK / 32];
    const ViewPtrs& v = blockIdx.y ? v1 : v0;
    const int line = blockIdx.x, Rp = dm.Rp, tid = threadIdx.x;
    const size_t estride = VERT ? (size_t)dm.W * Rp : (size_t)Rp;  // floats between consecutive positions of the line
    float* first = v.vol.tail + (VERT ? (size_t)line * Rp : (size_t)line * dm.W * Rp) + 2 * blockIdx.z;
    for (int x = tid; x < len; x += AGT_BLOCK) {
        const float2 c = *reinterpret_cast<const float2*>(first + (size_t)x * estride);
        P[x + 1] = make_double2((double)c.x, (double)c.y);
    }
    if (tid == 0) { P[0] = make_double2(0.0, 0.0); if (MODE == 2) Q[0] = make_double2(0.0, 0.0); }
    __syncthreads();
    tail_scan(P, warp_tot, len);
    const uint32_t* desc = VERT ? v.desc_v + (size_t)line * dm.Hd() : v.desc_h + (size_t)line * dm.Wd();
    for (int x = tid; x < len; x += AGT_BLOCK) {
        const uint32_t w = desc[x];
        const int a = w & 0xff, b = (w >> 8) & 0xff;
        const double2 hi = P[x + b + 1], lo = P[x - a];
        float r0 = __double2float_rn(hi.x - lo.x), r1 = __double2float_rn(hi.y - lo.y);
        if (MODE != 0) {
            const RcpN rn = rcp_prepare((float)(w >> 16));
            r0 = div_exact(r0, rn);
            r1 = div_exact(r1, rn);
        }
        if (MODE == 2) Q[x + 1] = make_double2((double)r0, (double)r1);
        else *reinterpret_cast<float2*>(first + (size_t)x * estride) = make_float2(r0, r1);
    }
    if (MODE == 2) {
        __syncthreads();
        tail_scan(Q, warp_tot, len);
        for (int x = tid; x < len; x += AGT_BLOCK) {
            const uint32_t w = desc[x];
            const int a = w & 0xff, b = (w >> 8) & 0xff;
            const double2 hi = Q[x + b + 1], lo = Q[x - a];
            *reinterpret_cast<float2*>(first + (size_t)x * estride) =
                make_float2(__double2float_rn(hi.x - lo.x), __double2float_rn(hi.y - lo.y));
        }
    }
}

template <bool VERT, int MODE>
static void launch_tail(const Launcher& L, const Dims& d, const ViewPtrs& left, const ViewPtrs& right)
{
    if (d.tail() <= 0) return;
    const int len = VERT ? d.H : d.W, nlines = VERT ? d.W : d.H;
    const size_t smem = (size_t)(len + 1) * sizeof(double2) * (MODE == 2 ? 2 : 1);
    static PerDevice smem_set;
    if (smem > 48 * 1024 && smem > smem_set.cur()) {
        cudaFuncSetAttribute(k_agg_tail<VERT, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        smem_set.cur() = smem;
    }
    dim3 grid((unsigned)nlines, 2, (unsigned)((d.tail() + 1) / 2));
    k_agg_tail<VERT, MODE><<<grid, AGT_BLOCK, smem, L.stream>>>(d, left, right);
    L.count(1);
}

// Generic guarded variant for lines shorter than AGG_LAG + 1 + AGG_PF (tiny images): one
// chain per thread, 68-entry ring, no prefetch.  Same arithmetic as k_agg_walk.
constexpr int AGS_BLOCK = 128, AGS_RING = 2 * kMaxArm + 2;
template <bool VERT, bool NORM>
__global__ void __launch_bounds__(AGS_BLOCK)
k_agg_small(Dims dm, ViewPtrs v0, ViewPtrs v1, int nb_main)
{
    extern __shared__ double sring[];  // [AGS_RING][AGS_BLOCK]
    const ViewPtrs& v = blockIdx.y ? v1 : v0;
    const int H = dm.H, W = dm.W;
    const int nlines = VERT ? W : H, len = VERT ? H : W;
    float* part;
    int pitch, nd;
    long long chain;
    if ((int)blockIdx.x < nb_main) {
        part = v.vol.main; pitch = dm.Dm; nd = dm.Dm;
        chain = (long long)blockIdx.x * AGS_BLOCK + threadIdx.x;
    } else {
        part = v.vol.tail; pitch = dm.Rp; nd = dm.tail();
        chain = (long long)((int)blockIdx.x - nb_main) * AGS_BLOCK + threadIdx.x;
    }
    if (chain >= (long long)nlines * nd) return;
    const int line = (int)(chain / nd), d = (int)(chain % nd);
    const size_t line_px = VERT ? (size_t)line : (size_t)line * W;
    float* cell = part + line_px * pitch + d;
    const size_t cstride = VERT ? (size_t)W * pitch : (size_t)pitch;
    const uint32_t* desc = VERT ? v.desc_v + (size_t)line * dm.Hd() : v.desc_h + (size_t)line * dm.Wd();
    double* my = sring + threadIdx.x;
    my[0] = 0.0;
    double P = 0.0;
    int head = 0;
    for (int t = 0; t < len + AGG_LAG; ++t) {
        const int o = t - AGG_LAG;
        if (t < len) {
            P += (double)cell[(size_t)t * cstride];
            head = (head + 1 == AGS_RING) ? 0 : head + 1;
            my[head * AGS_BLOCK] = P;
        }
        if (o >= 0 && o < len) {
            const uint32_t w = desc[o];
            const int a = w & 0xff, b = (w >> 8) & 0xff;
            const int newest = (t < len) ? t + 1 : len;
            int s1 = head - (newest - (o + b + 1)), s0 = head - (newest - (o - a));
            s1 += (s1 < 0) ? AGS_RING : 0;
            s0 += (s0 < 0) ? AGS_RING : 0;
            float r = __double2float_rn(my[s1 * AGS_BLOCK] - my[s0 * AGS_BLOCK]);
            if (NORM) r = div_exact(r, rcp_prepare((float)(w >> 16)));
            cell[(size_t)o * cstride] = r;
        }
    }
}

// Ring-home mix of the persistent kernel: tm TMEM warps + sm shared-memory warps per SM.
// TSM_AGG_MIX="tmH,smH,tmV,smV" overrides it (kernel experiments), "0" selects the shared-memory-only kernel.
struct AggMix {
    int tm[2], sm[2];  // [0] horizontal, [1] vertical pass
    bool persist;
};
static const AggMix& agg_mix()
{
    static const AggMix m = [] {
        AggMix r = {{4, 4}, {6, 6}, true};  // measured best at 1080p D=192 (profiles/README.md)
        if (const char* e = getenv("TSM_AGG_MIX")) {
            int a, b, c, d;
            if (sscanf(e, "%d,%d,%d,%d", &a, &b, &c, &d) == 4 && a >= 0 && a <= AGH_TM_MAX && c >= 0 && c <= AGH_TM_MAX &&
                b >= 1 && b <= AGH_SM_WARPS && d >= 1 && d <= AGH_SM_WARPS) {
                r.tm[0] = a; r.sm[0] = b; r.tm[1] = c; r.sm[1] = d;
            } else {
                r.persist = false;
            }
        }
        return r;
    }();
    return m;
}

// The tail part's kernels go to the context's side stream when the caller forked one (aggregate(), fused path): every
// disparity plane is aggregated independently, so the tail chain only depends on itself.
static Launcher tail_launcher(const Launcher& L)
{
    Launcher t = L;
    if (L.side) { t.stream = L.side; t.mark = nullptr; }
    return t;
}

// Experiment (TSM_AGG_SPARE_SMS=k): the persistent aggregation kernels take n_sm - k SMs (their warps pull work items from a
// global counter, so any grid is correct) and leave k SMs to the small kernels of the other pairs in flight.  Measured at C3
// with six pairs in flight, same call: profiles/README.md.  Default 0.
static int agg_spare_sms()
{
    static const int k = [] {
        const char* e = getenv("TSM_AGG_SPARE_SMS");
        const int v = e ? atoi(e) : 0;
        return v < 0 ? 0 : (v > 64 ? 64 : v);
    }();
    return k;
}

template <bool VERT, bool NORM>
static void launch_walk(const Launcher& L, const Dims& d, const ViewPtrs& left, const ViewPtrs& right, unsigned* ctr)
{
    const int len = VERT ? d.H : d.W;
    const long long nl = VERT ? d.W : d.H;
    const AggMix& mix = agg_mix();
    if (!VERT && len >= AGG_LAG + 1 + AGG_U) launch_tail<false, NORM ? 1 : 0>(tail_launcher(L), d, left, right);
    if (len >= AGG_LAG + 1 + AGG_U && AGG_NC == 2 && d.Dm >= 64 && d.Dm % 64 == 0 && mix.persist) {
        static PerDevice sm_count;  // doubles as "attribute set on this device"
        if (!sm_count.cur()) {
            cudaFuncSetAttribute(k_agg_persist<VERT, NORM>, cudaFuncAttributeMaxDynamicSharedMemorySize, AGH_SMEM);
            int dev = 0, n = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
            sm_count.cur() = (size_t)n;
        }
        const int n_sm = (int)sm_count.cur() - agg_spare_sms();
        const int tm = mix.tm[VERT], sm = mix.sm[VERT];
        k_agg_persist<VERT, NORM><<<n_sm, 32 * (tm + sm), AGH_SMEM, L.stream>>>(d, left, right, tm, ctr);
    } else if (len >= AGG_LAG + 1 + AGG_U) {
        static PerDevice attr_set;
        if (!attr_set.cur()) {
            cudaFuncSetAttribute(k_agg_walk<VERT, NORM>, cudaFuncAttributeMaxDynamicSharedMemorySize, SmemRing<AGG_BLOCK>::SPAN);
            attr_set.cur() = 1;
        }
        const int nb_main = (int)((nl * (d.Dm / AGG_NC) + AGG_BLOCK - 1) / AGG_BLOCK);
        const int nb_tail = VERT ? (int)((nl * ((d.tail() + AGG_NC - 1) / AGG_NC) + AGG_BLOCK - 1) / AGG_BLOCK) : 0;
        dim3 grid((unsigned)(nb_main + nb_tail), 2);
        if (grid.x == 0) return;  // Dn < 32, horizontal pass: the tail kernel did all of it
        k_agg_walk<VERT, NORM><<<grid, AGG_BLOCK, SmemRing<AGG_BLOCK>::SPAN, L.stream>>>(d, left, right, nb_main);
    } else {
        static PerDevice attr_set;
        const size_t smem = (size_t)AGS_RING * AGS_BLOCK * sizeof(double);
        if (!attr_set.cur()) {
            cudaFuncSetAttribute(k_agg_small<VERT, NORM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            attr_set.cur() = 1;
        }
        const int nb_main = (int)((nl * d.Dm + AGS_BLOCK - 1) / AGS_BLOCK);
        const int nb_tail = (int)((nl * d.tail() + AGS_BLOCK - 1) / AGS_BLOCK);
        dim3 grid((unsigned)(nb_main + nb_tail), 2);
        k_agg_small<VERT, NORM><<<grid, AGS_BLOCK, smem, L.stream>>>(d, left, right, nb_main);
    }
    L.count(1);
}

size_t aggregate_overread_floats(const Dims& d)
{
    // k_agg_walk prefetches up to AGG_NBUF*AGG_PF positions past the end of a line; the vertical pass
    // therefore touches that many rows behind each part of the volume (pitch <= max(Dm, Rp)).
    return (size_t)(AGG_U + 1) * d.W * (size_t)(d.Dm > d.Rp ? d.Dm : d.Rp);
}

template <bool VERT>
static void launch_fused(const Launcher& L, const Dims& d, const ViewPtrs& left, const ViewPtrs& right, unsigned* ctr)
{
    static PerDevice sm_count;  // doubles as "attribute set on this device"
    if (!sm_count.cur()) {
        cudaFuncSetAttribute(k_agg_fused<VERT>, cudaFuncAttributeMaxDynamicSharedMemorySize, AGF_SMEM);
        int dev = 0, n = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        sm_count.cur() = (size_t)n;
    }
    launch_tail<VERT, 2>(tail_launcher(L), d, left, right);  // the tail part's share: staged lines, both passes in shared memory
    k_agg_fused<VERT><<<(int)sm_count.cur() - agg_spare_sms(), AGF_THREADS, AGF_SMEM, L.stream>>>(d, left, right, ctr);
    L.count(1);
}

// TSM_AGG_FUSE=0 keeps the eight single-pass launches (A/B measurements); default: H | V.n+V | H.n+H | V.n+V | H.n
static bool agg_fuse_enabled()
{
    static const bool on = [] {
        const char* e = getenv("TSM_AGG_FUSE");
        return !(e && e[0] == '0');
    }();
    return on;
}

void aggregate(const Launcher& L, const Dims& d, const ViewPtrs& left, const ViewPtrs& right, unsigned* work_counters)
{
    // two work counters (main items, tail items) per pass of the persistent kernels
    cudaMemsetAsync(work_counters, 0, kAggCounterBytes, L.stream);
    if (agg_fuse_enabled() && AGG_NC == 2 && d.Dm >= 64 && d.Dm % 64 == 0 && d.W >= AGG_LAG + 1 + AGG_U && agg_mix().persist) {
        // fork: the five tail-part launches run on the side stream next to the five persistent main-part kernels
        // (0.40 ms of small launches per 1080p pair when they were serialised in between, for 0.5 % of the cells)
        Launcher Lf = L;
        const bool fork = L.side && L.events && d.tail() > 0 && !getenv("TSM_AGG_NO_FORK");
        if (fork) {
            cudaEventRecord(L.events[0], L.stream);
            cudaStreamWaitEvent(L.side, L.events[0], 0);
        } else {
            Lf.side = nullptr;
        }
        const Launcher& L = Lf;  // NOLINT: shadows the parameter on purpose
        unsigned* ctr = work_counters;
        L.begin("aggregate/h");
        launch_walk<false, false>(L, d, left, right, ctr);
        L.end();
        for (int k = 0; k < kIterations - 1; ++k) {  // iterations alternate H,V | V,H: the pass ending one and the pass starting the next share a direction
            ctr += 2;
            if (k % 2 == 0) {
                L.begin("aggregate/v_norm+v");
                launch_fused<true>(L, d, left, right, ctr);
            } else {
                L.begin("aggregate/h_norm+h");
                launch_fused<false>(L, d, left, right, ctr);
            }
            L.end();
        }
        static_assert(kIterations % 2 == 0, "the last pass is a horizontal normalising one");
        ctr += 2;
        L.begin("aggregate/h_norm");
        launch_walk<false, true>(L, d, left, right, ctr);
        L.end();
        if (fork) {  // join
            cudaEventRecord(L.events[1], L.side);
            cudaStreamWaitEvent(L.stream, L.events[1], 0);
        }
        return;
    }
    Launcher Ls = L;
    Ls.side = nullptr;  // the single-pass paths below keep everything on one stream
    const Launcher& Lserial = Ls;
    bool hf = true;
    unsigned* ctr = work_counters;
    for (int it = 0; it < kIterations; ++it, ctr += 4) {
        if (hf) {
            L.begin("aggregate/h");
            launch_walk<false, false>(Lserial, d, left, right, ctr);
            L.end();
            L.begin("aggregate/v_norm");
            launch_walk<true, true>(Lserial, d, left, right, ctr + 2);
            L.end();
        } else {
            L.begin("aggregate/v");
            launch_walk<true, false>(Lserial, d, left, right, ctr);
            L.end();
            L.begin("aggregate/h_norm");
            launch_walk<false, true>(Lserial, d, left, right, ctr + 2);
            L.end();
        }
        hf = !hf;
    }
}

// ---- self-test: div_exact == __fdiv_rn for every divisor the aggregation can see ----
__global__ void k_selftest_div(unsigned long long* mismatches)
{
    // blockIdx.x + 1 = divisor b in [1, 4489]; 2^21 dividends per divisor: a dense sweep of mantissas at the
    // magnitudes aggregation produces (sums in [0, 8978]) plus exact multiples of b
    const float b = (float)(blockIdx.x + 1);
    const RcpN rn = rcp_prepare(b);
    unsigned long long bad = 0;
    for (unsigned i = threadIdx.x; i < (1u << 21); i += blockDim.x) {
        float a;
        if (i & 1) a = __uint_as_float(0x3a000000u + i * 1021u);  // ~[4.9e-4, 1.3e4], stride-1021 mantissa walk
        else a = b * (float)(i >> 1) * 0.001953125f;                // multiples of b / 512
        if (!(a >= 0.f && a < 16384.f)) continue;
        const unsigned want = __float_as_uint(__fdiv_rn(a, b));
        if (__float_as_uint(div_exact(a, rn)) != want) ++bad;
        if (__float_as_uint(div_exact_rn(a, b, __frcp_rn(b))) != want) ++bad;
    }
    if (bad) atomicAdd(mismatches, bad);
}

void selftest_div(const Launcher& L, unsigned long long* d_mismatches)
{
    k_selftest_div<<<4489, 256, 0, L.stream>>>(d_mismatches);
    L.count(1);
}

}  // namespace tsm
