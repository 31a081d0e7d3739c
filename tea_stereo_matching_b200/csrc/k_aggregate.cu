// k_aggregate.cu -- cross-based cost aggregation (costAggregate, reference
// source/ADCensus.cpp:685-793), 4 iterations, pass order H,V | V,H | H,V | V,H.
//
// One aggregation1D pass is  out(p) = sum_{j=-a(p)}^{b(p)} in(p + j*r)  along a line
// (ADCensus.cpp:711-718); after the second pass of an iteration the plane is divided
// by the cross-window size (:743-749).  The reference adds in fp32 in ascending order;
// here every (line, d) chain walks its line once, keeps a running fp64 prefix sum P and
// emits  P[o+b+1] - P[o-a]  with a lag of 33 (= max arm); the last 72 prefixes live in a
// shared-memory ring.  fp64 prefix differences stay within ~4e-7 relative of the
// reference's sequential fp32 sums after all 4 iterations (fp32 prefixes do not;
// SURVEY 0.6).  Each cell is read once and written once per pass, in place.
//
// Thread mapping (k_agg_walk): one thread owns TWO adjacent disparities (d, d+1) of one
// line: 8-byte global accesses, 16-byte ring entries, and the per-step index arithmetic
// (ring slots, arm decode, pointers) is paid once per two cells.  Lanes run over d, so a
// warp reads 256 contiguous bytes per step.  Inputs do not depend on the recurrence:
// they are loaded AGG_PF steps ahead into a register ring, so every warp keeps AGG_PF
// 256-byte loads in flight and the only loop-carried dependency is the fp64 prefix add.
// The main loop is unrolled AGG_PF times and the ring length is a multiple of AGG_PF,
// so ring slots of the pushes are compile-time offsets and the wrap test runs once per
// AGG_PF steps.
//
// Division: the reference divides the fp32 sum by (float)N with an IEEE fp32 divide
// (ADCensus.cpp:747); div_exact(fl32(sum), N) (tsm_common.cuh) is that divide, correctly rounded,
// without the slow-path branch of __fdiv_rn.
//
// Per-position side data (the two arm lengths of the pass and N) come as ONE 32-bit step
// descriptor (k_prep.cu), four consecutive positions per 16-byte load: small broadcast loads
// cost real bandwidth on this access pattern (measured: -15 % with two scalar loads per step).
#include "tsm_common.cuh"

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
#ifndef TSM_AGG_NBUF
#define TSM_AGG_NBUF 8
#endif
constexpr int AGG_NBUF = TSM_AGG_NBUF;    // register buffers: NBUF-1 batches in flight
constexpr int AGG_U = AGG_NBUF * AGG_PF;  // steps per main-loop iteration
#ifndef TSM_AGG_SUB
#define TSM_AGG_SUB 4
#endif
constexpr int AGG_SUB = TSM_AGG_SUB;      // steps per software-pipelined sub-block of a batch
static_assert(AGG_PF % AGG_SUB == 0, "sub-blocks tile a batch");
constexpr int AGG_RING = 72;              // >= 68 prefixes, multiple of AGG_PF
constexpr int AGG_SLOT = AGG_BLOCK * 8 * AGG_NC;  // bytes between consecutive ring slots
constexpr int AGG_RING_BYTES = AGG_RING * AGG_SLOT;
// P[i] lives in slot (i + AGG_P0) mod AGG_RING, chosen so that P[AGG_LAG + 1] (the first
// prefix written in the main loop) sits on a multiple of AGG_PF.
constexpr int AGG_P0 = (AGG_PF - (AGG_LAG + 1) % AGG_PF) % AGG_PF;
static_assert(AGG_RING % AGG_PF == 0 && AGG_RING >= 2 * kMaxArm + 2, "ring geometry");
static_assert(AGG_PF <= AGG_LAG, "arm prefetch must stay inside the line");

__device__ __forceinline__ void st_ring(uint32_t addr, double a, double b)
{
    if (AGG_NC == 2) asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(a), "d"(b) : "memory");
    else asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(a) : "memory");
}
__device__ __forceinline__ void ld_ring(uint32_t addr, double& a, double& b)
{
    if (AGG_NC == 2) asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "r"(addr) : "memory");
    else { asm volatile("ld.shared.f64 %0, [%1];" : "=d"(a) : "r"(addr) : "memory"); b = 0.0; }
}

// Streaming accesses of the cost volume: every cell is touched exactly once per pass, so the
// lines must not occupy the (small, shared-memory-carved) L1.
__device__ __forceinline__ float2 ld_stream(const float* p)
{
    float2 v;
    if (AGG_NC == 2) asm volatile("ld.global.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    else { asm volatile("ld.global.L1::no_allocate.f32 %0, [%1];" : "=f"(v.x) : "l"(p)); v.y = 0.f; }
    return v;
}
__device__ __forceinline__ void st_stream(float* p, float a, float b)
{
    if (AGG_NC == 2) asm volatile("st.global.L1::no_allocate.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
    else asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(p), "f"(a) : "memory");
}

static_assert(AGG_PF % 4 == 0, "descriptors are fetched four at a time; batches run in sub-blocks of 4");

// Requires len >= AGG_LAG + 1 (host-checked) and AGG_PF positions of over-read slack
// behind every line end (the volumes are allocated with it).
template <bool VERT, bool NORM>
__global__ void __launch_bounds__(AGG_BLOCK)
k_agg_walk(Dims dm, ViewPtrs v0, ViewPtrs v1, int wsel, int nb_main)
{
    extern __shared__ __align__(16) unsigned char ring_raw[];  // [AGG_RING][AGG_BLOCK] x double2
    const ViewPtrs& v = blockIdx.y ? v1 : v0;
    const int H = dm.H, W = dm.W;
    const int nlines = VERT ? W : H, len = VERT ? H : W;
    // chains of the 128-byte aligned main part first, then (blocks >= nb_main) those of the tail part
    float* part;
    int pitch, npair;
    long long chain;
    if ((int)blockIdx.x < nb_main) {
        part = v.vol.main; pitch = dm.Dm; npair = dm.Dm / AGG_NC;
        chain = (long long)blockIdx.x * AGG_BLOCK + threadIdx.x;
    } else {
        part = v.vol.tail; pitch = dm.Rp; npair = (dm.tail() + AGG_NC - 1) / AGG_NC;
        chain = (long long)((int)blockIdx.x - nb_main) * AGG_BLOCK + threadIdx.x;
    }
    if (chain >= (long long)nlines * npair) return;
    const int line = (int)(chain / npair), d = AGG_NC * (int)(chain % npair);

    const size_t cstride = VERT ? (size_t)W * pitch : (size_t)pitch;  // floats between consecutive positions
    const size_t line_px = VERT ? (size_t)line : (size_t)line * W;
    const float* in_ptr = part + line_px * pitch + d;
    float* out_ptr = part + line_px * pitch + d;
    const uint32_t* desc_line = VERT ? v.desc_v + (size_t)line * dm.Hd() : v.desc_h + (size_t)line * dm.Wd();
    const uint32_t* desc_ptr = desc_line;

    const uint32_t ring0 = (uint32_t)__cvta_generic_to_shared(ring_raw) + threadIdx.x * (8 * AGG_NC);
    double P0 = 0.0, P1 = 0.0;
    st_ring(ring0 + AGG_P0 * AGG_SLOT, 0.0, 0.0);  // P[0]

    // output o, given the (possibly virtual) slot byte offset `top` of P[o + 34]
    auto output = [&](uint32_t top, uint32_t desc) {
        const int a = desc & 0xff, b = (desc >> 8) & 0xff;
        int s1 = (int)top - (AGG_LAG - b) * AGG_SLOT;      // P[o + b + 1]
        int s0 = (int)top - (AGG_LAG + 1 + a) * AGG_SLOT;  // P[o - a]
        s1 += (s1 < 0) ? AGG_RING_BYTES : 0;
        s0 += (s0 < 0) ? AGG_RING_BYTES : 0;
        double h0, h1, l0, l1;
        ld_ring(ring0 + s1, h0, h1);
        ld_ring(ring0 + s0, l0, l1);
        float r0 = __double2float_rn(h0 - l0), r1 = __double2float_rn(h1 - l1);
        if (NORM) {
            const RcpN rn = rcp_prepare((float)(desc >> 16));
            r0 = div_exact(r0, rn);
            r1 = div_exact(r1, rn);
        }
        st_stream(out_ptr, r0, r1);
        out_ptr += cstride;
    };

    // ---- fill: pushes t = 0 .. AGG_LAG-1 (P[1..33] -> slots AGG_P0+1 .. AGG_P0+33, no wrap),
    // values fetched in 3 batches of 11 ----
    uint32_t hs = AGG_P0 * AGG_SLOT;  // slot of the newest prefix
    static_assert(AGG_LAG == 33, "fill batches");
#pragma unroll 1
    for (int g = 0; g < 3; ++g) {
        float2 tmp[11];
#pragma unroll
        for (int u = 0; u < 11; ++u) tmp[u] = ld_stream(in_ptr + (size_t)u * cstride);
        in_ptr += (size_t)11 * cstride;
#pragma unroll
        for (int u = 0; u < 11; ++u) {
            P0 += (double)tmp[u].x;
            P1 += (double)tmp[u].y;
            hs += AGG_SLOT;
            st_ring(ring0 + hs, P0, P1);
        }
    }
    // newest = P[33] at slot AGG_P0 + 33; the next prefix P[34] goes to a multiple of AGG_PF.
    uint32_t nx = hs + AGG_SLOT;  // slot of the next prefix to be written, multiple of AGG_PF slots
    if (nx == AGG_RING_BYTES) nx = 0;

    // ---- main: steps t = AGG_LAG .. len-1: push in[t] -> P[t+1], emit o = t - AGG_LAG ----
    // Software pipeline in BATCHES of AGG_PF steps with two register buffers: all loads of batch
    // i+1 are issued back to back, then batch i (whose loads were issued one batch earlier) is
    // processed.  Batching matters: a warp has only six scoreboard slots and a slot completes
    // when ALL loads charged to it have landed, so independent loads must be grouped by the
    // time they are needed, not interleaved one per step.
    float2 vin[AGG_NBUF][AGG_PF];
    uint32_t av[AGG_NBUF][AGG_PF];
    auto load_batch = [&](int buf) {
#pragma unroll
        for (int u = 0; u < AGG_PF; ++u) vin[buf][u] = ld_stream(in_ptr + (size_t)u * cstride);
#pragma unroll
        for (int j = 0; j < AGG_PF / 4; ++j) {
            const uint4 q = *reinterpret_cast<const uint4*>(desc_ptr + 4 * j);
            av[buf][4 * j + 0] = q.x; av[buf][4 * j + 1] = q.y; av[buf][4 * j + 2] = q.z; av[buf][4 * j + 3] = q.w;
        }
        in_ptr += (size_t)AGG_PF * cstride;
        desc_ptr += AGG_PF;
    };
    // One batch = AGG_PF steps, processed in sub-blocks of 4: (A) the four prefix pushes and their ring
    // stores, (B) all eight ring loads, (C) the four outputs.  A warp issues in order, so finishing
    // step u right after its own ring loads would expose the LDS -> DADD -> F2F -> STG latency on every
    // step; grouped like this it is paid once per four steps.
    auto run_batch = [&](int buf, int nsteps) {  // nsteps == AGG_PF in the steady state
#pragma unroll
        for (int u0 = 0; u0 < AGG_PF; u0 += AGG_SUB) {
            double h0[AGG_SUB], h1[AGG_SUB], l0[AGG_SUB], l1[AGG_SUB];
            RcpN rn[AGG_SUB];
#pragma unroll
            for (int w = 0; w < AGG_SUB; ++w) {
                const int u = u0 + w;
                if (NORM && u < nsteps) rn[w] = rcp_prepare((float)(av[buf][u] >> 16));  // off the critical path
                if (u < nsteps) {
                    P0 += (double)vin[buf][u].x;
                    P1 += (double)vin[buf][u].y;
                    st_ring(ring0 + nx + u * AGG_SLOT, P0, P1);
                }
            }
#pragma unroll
            for (int w = 0; w < AGG_SUB; ++w) {
                const int u = u0 + w;
                if (u < nsteps) {
                    const uint32_t desc = av[buf][u];
                    const int a = desc & 0xff, b = (desc >> 8) & 0xff;
                    const uint32_t top = nx + u * AGG_SLOT;
                    int s1 = (int)top - (AGG_LAG - b) * AGG_SLOT;      // P[o + b + 1]
                    int s0 = (int)top - (AGG_LAG + 1 + a) * AGG_SLOT;  // P[o - a]
                    s1 += (s1 < 0) ? AGG_RING_BYTES : 0;
                    s0 += (s0 < 0) ? AGG_RING_BYTES : 0;
                    ld_ring(ring0 + s1, h0[w], h1[w]);
                    ld_ring(ring0 + s0, l0[w], l1[w]);
                }
            }
#pragma unroll
            for (int w = 0; w < AGG_SUB; ++w) {
                const int u = u0 + w;
                if (u < nsteps) {
                    float r0 = __double2float_rn(h0[w] - l0[w]), r1 = __double2float_rn(h1[w] - l1[w]);
                    if (NORM) {
                        r0 = div_exact(r0, rn[w]);
                        r1 = div_exact(r1, rn[w]);
                    }
                    st_stream(out_ptr, r0, r1);
                    out_ptr += cstride;
                }
            }
        }
    };
    const int nB = len - AGG_LAG;
    uint32_t newest = hs;  // slot of the newest prefix
    int done = 0;
    // AGG_NBUF register buffers: NBUF-1 batches are in flight while one is being processed.
#pragma unroll
    for (int b = 0; b < AGG_NBUF - 1; ++b) load_batch(b);
    for (; done + AGG_NBUF * AGG_PF <= nB; done += AGG_NBUF * AGG_PF) {
#pragma unroll
        for (int b = 0; b < AGG_NBUF; ++b) {
            load_batch((b + AGG_NBUF - 1) % AGG_NBUF);
            run_batch(b, AGG_PF);
            nx += AGG_PF * AGG_SLOT;
            if (nx == AGG_RING_BYTES) nx = 0;
        }
    }
    // tail: fewer than NBUF*PF steps left; buffers 0 .. NBUF-2 already hold the next batches
    {
        int rem = nB - done;
        newest = (nx == 0 ? AGG_RING_BYTES : nx) - AGG_SLOT;
#pragma unroll
        for (int b = 0; b < AGG_NBUF; ++b) {
            if (rem > 0) {
                if (b == AGG_NBUF - 1) load_batch(b);  // the one buffer that was not prefetched
                const int n = rem < AGG_PF ? rem : AGG_PF;
                run_batch(b, n);
                newest = nx + (uint32_t)(n - 1) * AGG_SLOT;
                nx += AGG_PF * AGG_SLOT;
                if (nx == AGG_RING_BYTES) nx = 0;
                rem -= n;
            }
        }
    }
    // ---- drain: o = len-LAG .. len-1.  Newest prefix stays P[len]; the virtual slot of P[o+34]
    // is (j+1) slots past it.  Only real slots (<= o+b+1 <= len) are read.
    for (int j = 0; j < AGG_LAG; ++j) {
        const int o = len - AGG_LAG + j;
        uint32_t top = newest + (uint32_t)(j + 1) * AGG_SLOT;
        top -= (top >= AGG_RING_BYTES) ? AGG_RING_BYTES : 0;
        output(top, desc_line[o]);
    }
}

// Generic guarded variant for lines shorter than AGG_LAG + 1 + AGG_PF (tiny images): one
// chain per thread, 68-entry ring, no prefetch.  Same arithmetic as k_agg_walk.
constexpr int AGS_BLOCK = 128, AGS_RING = 2 * kMaxArm + 2;
template <bool VERT, bool NORM>
__global__ void __launch_bounds__(AGS_BLOCK)
k_agg_small(Dims dm, ViewPtrs v0, ViewPtrs v1, int wsel, int nb_main)
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

template <bool VERT, bool NORM>
static void launch_walk(const Launcher& L, const Dims& d, const ViewPtrs& left, const ViewPtrs& right, int wsel)
{
    const int len = VERT ? d.H : d.W;
    if (len >= AGG_LAG + 1 + AGG_U) {
        static bool attr_set = false;
        if (!attr_set) {
            cudaFuncSetAttribute(k_agg_walk<VERT, NORM>, cudaFuncAttributeMaxDynamicSharedMemorySize, AGG_RING_BYTES);
            attr_set = true;
        }
        const long long nl = VERT ? d.W : d.H;
        const int nb_main = (int)((nl * (d.Dm / AGG_NC) + AGG_BLOCK - 1) / AGG_BLOCK);
        const int nb_tail = (int)((nl * ((d.tail() + AGG_NC - 1) / AGG_NC) + AGG_BLOCK - 1) / AGG_BLOCK);
        dim3 grid((unsigned)(nb_main + nb_tail), 2);
        k_agg_walk<VERT, NORM><<<grid, AGG_BLOCK, AGG_RING_BYTES, L.stream>>>(d, left, right, wsel, nb_main);
    } else {
        static bool attr_set = false;
        const size_t smem = (size_t)AGS_RING * AGS_BLOCK * sizeof(double);
        if (!attr_set) {
            cudaFuncSetAttribute(k_agg_small<VERT, NORM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            attr_set = true;
        }
        const long long nl = VERT ? d.W : d.H;
        const int nb_main = (int)((nl * d.Dm + AGS_BLOCK - 1) / AGS_BLOCK);
        const int nb_tail = (int)((nl * d.tail() + AGS_BLOCK - 1) / AGS_BLOCK);
        dim3 grid((unsigned)(nb_main + nb_tail), 2);
        k_agg_small<VERT, NORM><<<grid, AGS_BLOCK, smem, L.stream>>>(d, left, right, wsel, nb_main);
    }
    L.count(1);
}

size_t aggregate_overread_floats(const Dims& d)
{
    // k_agg_walk prefetches up to AGG_NBUF*AGG_PF positions past the end of a line; the vertical pass
    // therefore touches that many rows behind each part of the volume (pitch <= max(Dm, Rp)).
    return (size_t)(AGG_U + 1) * d.W * (size_t)(d.Dm > d.Rp ? d.Dm : d.Rp);
}

void aggregate(const Launcher& L, const Dims& d, const ViewPtrs& left, const ViewPtrs& right)
{
    bool hf = true;
    for (int it = 0; it < kIterations; ++it) {
        if (hf) {
            launch_walk<false, false>(L, d, left, right, 0);
            launch_walk<true, true>(L, d, left, right, 0);
        } else {
            launch_walk<true, false>(L, d, left, right, 1);
            launch_walk<false, true>(L, d, left, right, 1);
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
        if (__float_as_uint(div_exact(a, rn)) != __float_as_uint(__fdiv_rn(a, b))) ++bad;
    }
    if (bad) atomicAdd(mismatches, bad);
}

void selftest_div(const Launcher& L, unsigned long long* d_mismatches)
{
    k_selftest_div<<<4489, 256, 0, L.stream>>>(d_mismatches);
    L.count(1);
}

}  // namespace tsm
