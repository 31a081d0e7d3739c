// k_aggregate.cu -- cross-based cost aggregation (costAggregate, reference
// source/ADCensus.cpp:685-793), 4 iterations, pass order H,V | V,H | H,V | V,H.
//
// One aggregation1D pass is  out(p) = sum_{j=-a(p)}^{b(p)} in(p + j*r)  along a line
// (ADCensus.cpp:711-718); after the second pass of an iteration the plane is divided
// by the cross-window size (:743-749).  The reference adds in fp32 in ascending order;
// here every (line, d) chain is one thread that walks the line once, keeps a running
// fp64 prefix sum P and emits  P[o+b+1] - P[o-a]  with a lag of 33 (= max arm), the
// last 68 prefixes living in a shared-memory ring.  fp64 prefix differences are within
// ~4e-7 relative of the reference's sequential fp32 sums after all 4 iterations
// (fp32 prefixes are not).  Each cell is read once and written once per
// pass, in place, with lanes over d (coalesced 128-byte rows).
//
// The division is the reference's own IEEE fp32 divide of the fp32-rounded sum by
// (float)N (ADCensus.cpp:747).
#include "tsm_common.cuh"

namespace tsm {

constexpr int AGG_BLOCK = 128;
constexpr int AGG_LAG = kMaxArm;       // 33
constexpr int AGG_RING = 2 * kMaxArm + 2;  // 68 prefixes: P[o-33] .. P[o+34]
constexpr int AGG_U = 4;               // steps per batch (loads issued together)

template <bool VERT, bool NORM>
__global__ void __launch_bounds__(AGG_BLOCK)
k_agg_walk(Dims dm, ViewPtrs v0, ViewPtrs v1, int wsel)
{
    extern __shared__ double ring[];  // [AGG_RING][AGG_BLOCK]
    const ViewPtrs& v = blockIdx.y ? v1 : v0;
    const int H = dm.H, W = dm.W, Dn = dm.Dn, Dp = dm.Dp;
    const int nlines = VERT ? W : H, len = VERT ? H : W;
    const long long chain = (long long)blockIdx.x * AGG_BLOCK + threadIdx.x;
    if (chain >= (long long)nlines * Dn) return;
    const int line = (int)(chain / Dn), d = (int)(chain % Dn);

    float* __restrict__ cell = v.vol + (VERT ? (size_t)line * Dp : (size_t)line * W * Dp) + d;
    const size_t cstride = VERT ? (size_t)W * Dp : (size_t)Dp;
    const uchar4* __restrict__ arm = v.arms + (VERT ? (size_t)line : (size_t)line * W);
    const size_t astride = VERT ? (size_t)W : 1;
    const float* __restrict__ wn = v.wsize + (size_t)wsel * H * W + (VERT ? (size_t)line : (size_t)line * W);

    double* my = ring + threadIdx.x;
    my[0] = 0.0;  // P[0]
    double P = 0.0;
    int head = 0;  // slot of P[t] at the start of step t

    for (int t0 = 0; t0 < len + AGG_LAG; t0 += AGG_U) {
        float vin[AGG_U];
        uchar4 av[AGG_U];
        float nn[AGG_U];
#pragma unroll
        for (int u = 0; u < AGG_U; ++u) {
            const int t = t0 + u, o = t - AGG_LAG;
            vin[u] = (t < len) ? cell[(size_t)t * cstride] : 0.f;
            av[u] = (o >= 0 && o < len) ? arm[(size_t)o * astride] : make_uchar4(0, 0, 0, 0);
            if (NORM) nn[u] = (o >= 0 && o < len) ? wn[(size_t)o * astride] : 1.f;
        }
#pragma unroll
        for (int u = 0; u < AGG_U; ++u) {
            const int t = t0 + u, o = t - AGG_LAG;
            if (t < len) {
                P += (double)vin[u];
                head = (head + 1 == AGG_RING) ? 0 : head + 1;  // slot of P[t+1]
                my[head * AGG_BLOCK] = P;
            }
            if (o >= 0 && o < len) {
                const int a = VERT ? av[u].x : av[u].z, b = VERT ? av[u].y : av[u].w;
                // newest stored prefix is P[min(t+1,len)] at slot `head`
                const int newest = (t < len) ? t + 1 : len;
                int s1 = head - (newest - (o + b + 1));
                int s0 = head - (newest - (o - a));
                s1 += (s1 < 0) ? AGG_RING : 0;
                s0 += (s0 < 0) ? AGG_RING : 0;
                const double sum = my[s1 * AGG_BLOCK] - my[s0 * AGG_BLOCK];
                float r = __double2float_rn(sum);
                if (NORM) r = __fdiv_rn(r, nn[u]);
                cell[(size_t)o * cstride] = r;
            }
        }
    }
}

template <bool VERT, bool NORM>
static void launch_walk(const Launcher& L, const Dims& d, const ViewPtrs& left, const ViewPtrs& right, int wsel)
{
    static bool attr_set = false;
    const size_t smem = (size_t)AGG_RING * AGG_BLOCK * sizeof(double);
    if (!attr_set) {
        cudaFuncSetAttribute(k_agg_walk<VERT, NORM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_set = true;
    }
    const long long chains = (long long)(VERT ? d.W : d.H) * d.Dn;
    dim3 grid((unsigned)((chains + AGG_BLOCK - 1) / AGG_BLOCK), 2);
    k_agg_walk<VERT, NORM><<<grid, AGG_BLOCK, smem, L.stream>>>(d, left, right, wsel);
    L.count(1);
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

}  // namespace tsm
