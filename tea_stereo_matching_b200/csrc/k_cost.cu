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
// Both views come from ONE evaluation: the cost depends only on the pair (xL, xR), so
//   C_1[d](y, c) = C_0[d](y, c + d)   for c + d < W,   2.f otherwise (xL leaves the image)
// i.e. the right volume is the left volume sheared along x.  The kernel computes the left tile and
// writes every value twice.
//
// Mapping: one CTA = (row y, 64 left pixels).  The signatures + BGRx pixels of the right view for
// the tile's whole disparity reach are staged once in shared memory.  A warp handles 4 consecutive
// left pixels at a time; lane l of chunk k owns ONE right column and evaluates it against the 4
// left pixels (disparities e, e+1, e+2, e+3), so each staged signature is read once per 4 cells.
// The six 32-bit match words of a cell are summed with a carry-save adder tree (3 POPC instead of
// 6: POPC is a quarter-rate instruction).  Costs land in a shared tile [64][Dn]; it is then
// flushed twice: as rows (left pixel vectors, aligned 16-byte stores) and along its diagonals
// (right pixel c holds tile[c + d - x0][d], a run of <= 64 consecutive d: 128-byte aligned warp
// stores; the diagonal walk has an odd bank stride, so it is conflict-free).
#include "tsm_common.cuh"
#include <type_traits>

namespace tsm {

#ifndef TSM_COST_WARPS
#define TSM_COST_WARPS 8
#endif
constexpr int COST_WARPS = TSM_COST_WARPS;
constexpr int COST_J = 4;       // fixed pixels per warp iteration
constexpr int TAB_C_N = kTabCensus;
constexpr int TAB_C_PAD_RGB = 192, TAB_C_PAD_HSI = 194;  // RGB: (768 + 192) * 4 = 3840 bytes, HSI: (2806 + 194) * 4 = 12000: what follows stays 16-byte aligned
constexpr float kAdSentinel = -4.f;  // "table value" of an invalid pixel pair: 2 - (-4) - exp >= 5, clamped to 2.f
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

#ifndef TSM_COST_MINB
#define TSM_COST_MINB 3
#endif
// COST_TX = left pixels per CTA: 64, or 32 when the shared tile [COST_TX][Dn] would otherwise leave one CTA per SM
// HSI: adCost = min(|dH|, 255 - |dH|) * 1 + |dS| * 2.5 + |dI| * 2.5 (computeHSIADCost, :439-452) is exact in fp32,
// so the table is indexed with 2 * adCost = 2 hd + 5 (ds + di) <= 2804.
// MASK (mask matching): a black centre pixel makes the census cost +inf, i.e. its exp() term 0 (:459, 481, 518); a
// cell whose OWN view's pixel is black is 2.f (:551-555) -- applied when the tile is flushed, per view.
template <int COST_TX, bool HSI, bool MASK>
__global__ void __launch_bounds__(COST_WARPS * 32, TSM_COST_MINB)
k_cost_init(Dims dm, ViewPtrs vl, ViewPtrs vr, const float* __restrict__ g_tab_ad, const float* __restrict__ g_tab_c)
{
    extern __shared__ __align__(16) uint32_t smem[];
    const int y = blockIdx.y, x0 = blockIdx.x * COST_TX;
    const int H = dm.H, W = dm.W, Dn = dm.Dn;
    const int ncol = (COST_TX + Dn - 1 + 3) & ~3;  // right columns the tile can reach (padded to keep 16-byte alignment)
    // the AD table + ONE sentinel entry behind it (index TAB_AD_USED) for pixel pairs that do not exist: a census window of the
    // right pixel outside the image adds a penalty to the AD sum, the index clamps to the sentinel, and the cost, clamped to 2.f from
    // above, is the reference's 2.f (:562-566).  A valid cell is 2 - exp(-ad/10) - exp(-census/30) <= 2 - exp(-189/30) < 2: the
    // clamp never touches it.  One FMNMX per cell instead of two selects on a combined predicate (the kernel is ALU-pipe bound).
    constexpr int TAB_AD_USED = HSI ? kTabAdHsi : kTabAdRgb;
    constexpr int TAB_AD_N = HSI ? kTabAdHsi + 1 : kTabAdRgb + 2;  // sentinel (+ padding): keeps the 16-byte alignment
    constexpr int TAB_C_PAD = HSI ? TAB_C_PAD_HSI : TAB_C_PAD_RGB;
    float* tab_ad = reinterpret_cast<float*>(smem);
    float* tab_c = tab_ad + TAB_AD_N;
    uint32_t* mw = reinterpret_cast<uint32_t*>(tab_c + TAB_C_PAD);  // [ncol][13]: 12 signature words + pixel (stride 13: conflict-free)
    for (int i = threadIdx.x; i < TAB_AD_USED; i += blockDim.x) tab_ad[i] = g_tab_ad[i];
    if (threadIdx.x == 0) tab_ad[TAB_AD_USED] = kAdSentinel;
    for (int i = threadIdx.x; i < TAB_C_N; i += blockDim.x) tab_c[i] = g_tab_c[i];

    const size_t npx = (size_t)H * W, row = (size_t)y * W;
    const int hw = kCensusW / 2;
    const bool yout = (y - kCensusH / 2 < 0) || (y + kCensusH / 2 >= H);
    const int cbase = x0 - (Dn - 1);  // first staged right column
    for (int i = threadIdx.x; i < ncol; i += blockDim.x) {
        const int c = cbase + i;
        const bool ok = !yout && c - hw >= 0 && c + hw < W;
        mw[13 * i + 12] = ok ? vr.img4[row + c] : kInvalidPix;
#pragma unroll
        for (int p = 0; p < 6; ++p) {
            const uint64_t s = ok ? vr.census[(size_t)p * npx + row + c] : 0ull;
            mw[13 * i + 2 * p] = (uint32_t)s;
            mw[13 * i + 2 * p + 1] = (uint32_t)(s >> 32);
        }
    }
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // lane-0 broadcast: provably warp-uniform for the compiler
    const int nchunk = (Dn + COST_J - 1 + 31) / 32;
    const int DnP = (Dn + 3) & ~3;
    float* ctile = reinterpret_cast<float*>(mw + 13 * ncol);  // [COST_TX][DnP], 16-byte aligned
    for (int g = warp; g < COST_TX / COST_J; g += COST_WARPS) {
        const int x = x0 + g * COST_J;
        if (x >= W) break;
        // the four left pixels x .. x+3 (uniform across the warp)
        Sig f[COST_J];
#pragma unroll
        for (int j = 0; j < COST_J; ++j) {
            const int xf = x + j;
            const bool ok = !yout && xf < W && xf - hw >= 0 && xf + hw < W;
            f[j].pix = ok ? vl.img4[row + xf] : kInvalidPix;
#pragma unroll
            for (int p = 0; p < 6; ++p) {
                const uint64_t s = ok ? vl.census[(size_t)p * npx + row + xf] : 0ull;
                f[j].w[2 * p] = (uint32_t)s;
                f[j].w[2 * p + 1] = (uint32_t)(s >> 32);
            }
        }
#pragma unroll
        for (int j = 0; j < COST_J; ++j) {
            // keep the left signatures in registers: without this the compiler re-loads all 28 words
            // from global memory in every chunk iteration instead of keeping 52 registers live
#pragma unroll
            for (int i = 0; i < 12; ++i) asm volatile("" : "+r"(f[j].w[i]));
            asm volatile("" : "+r"(f[j].pix));
        }
        float* tj[COST_J];  // tj[j] + e addresses disparity d_j = e + j of left pixel x + j
#pragma unroll
        for (int j = 0; j < COST_J; ++j) tj[j] = ctile + (g * COST_J + j) * DnP + j;
        // One chunk: 32 right columns against the four left pixels.  CHECK: the chunk can hold disparities outside [0, Dn) (the
        // first and the last chunks); the ones in between store without a predicate.
        auto chunk = [&](int k, auto check) {
            constexpr bool CHECK = decltype(check)::value;
            // e = disparity of this lane's right column c = x - e against left pixel j = 0; d_j = e + j
            const int e = 32 * k + lane - (COST_J - 1);
            const int ci = x - e - cbase;
            const bool cin = ci >= 0 && ci < ncol;
            const int cs = cin ? ci : 0;
            Sig m;
            const uint32_t* mc = mw + 13 * cs;
            m.pix = mc[12];
#pragma unroll
            for (int i = 0; i < 12; ++i) m.w[i] = mc[i];
            // a right pixel without a census window: the penalty pushes the table index to the sentinel
            const uint32_t pen = (cin && m.pix != kInvalidPix) ? 0u : 8192u;
#pragma unroll
            for (int j = 0; j < COST_J; ++j) {
                const int d = e + j;
                int ad3;
                if (HSI) {
                    const uint32_t dv = __vabsdiffu4(f[j].pix, m.pix);  // |dH|, |dS|, |dI| (byte 3 is 0 or 255 for invalid)
                    const int hd = dv & 0xffu;
                    ad3 = min(2 * min(hd, 255 - hd) + 5 * (int)(((dv >> 8) & 0xffu) + ((dv >> 16) & 0xffu)) + (int)pen, TAB_AD_USED);
                } else {
                    uint32_t sad;
                    asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(sad) : "r"(f[j].pix), "r"(m.pix), "r"(pen));
                    ad3 = min((int)sad, TAB_AD_USED);
                }
                const int cen = census_count(f[j], m);
                float ec = tab_c[cen];
                if (MASK && (f[j].pix == 0u || m.pix == 0u)) ec = 0.f;
                const float cost = fminf(__fsub_rn(__fsub_rn(2.f, tab_ad[ad3]), ec), 2.f);
                if (CHECK) {
                    // predicated store (as PTX: the compiler otherwise builds a reconvergence region around each of the four)
                    asm volatile("{ .reg .pred p; setp.lt.u32 p, %2, %3; @p st.shared.f32 [%0], %1; }" ::"r"(
                                     (uint32_t)__cvta_generic_to_shared(tj[j] + e)),
                                 "f"(cost), "r"((unsigned)d), "r"((unsigned)Dn)
                                 : "memory");
                } else {
                    tj[j][e] = cost;
                }
            }
        };
        // chunk k holds e = 32 k - 3 .. 32 k + 28, d = e .. e + 3: only chunk 0 can go below 0, only chunks with 32 k + 31 >= Dn above Dn - 1
        const int kfull = Dn / 32;  // chunks 1 .. kfull - 1 lie inside [0, Dn): 32 k + 31 <= Dn - 1
        chunk(0, std::true_type{});
        int k = 1;
        for (; k < kfull && k < nchunk; ++k) chunk(k, std::false_type{});
        for (; k < nchunk; ++k) chunk(k, std::true_type{});
    }
    __syncthreads();
    // Left pixels without a census window (the four border columns on either side; every pixel of the three border rows): the loop
    // above used an all-zero signature for them; their rows of the tile are 2.f (:562-566).  Only border tiles get here.
    {
        const int jlo = x0 == 0 ? hw : 0;                       // left pixels j < jlo are invalid
        const int jhi = yout ? 0 : min(COST_TX, W - hw - x0);   // left pixels j >= jhi are invalid
        if (jlo > 0 || jhi < COST_TX) {
            for (int j = 0; j < COST_TX; ++j) {
                if ((j < jlo || j >= jhi || yout) && x0 + j < W)
                    for (int d = threadIdx.x; d < Dn; d += blockDim.x) ctile[j * DnP + d] = 2.f;
            }
            __syncthreads();
        }
    }

    __syncthreads();

    const int Dm = dm.Dm, Rp = dm.Rp, r = dm.tail();
    const int npix = min(COST_TX, W - x0);
    // ---- left view: rows of the tile.  16-byte aligned vector stores of the main part, scalar stores of the tail part
    {
        const int nq = Dm / 4;  // float4 per pixel in the main part
        for (int j = warp; j < npix; j += COST_WARPS) {
            const float4* src = reinterpret_cast<const float4*>(ctile + j * DnP);
            float4* dst = reinterpret_cast<float4*>(vl.vol.main + (row + x0 + j) * Dm);
            const bool hole = MASK && vl.img4[row + x0 + j] == 0u;
            const float4 two = make_float4(2.f, 2.f, 2.f, 2.f);
            for (int q = lane; q < nq; q += 32) dst[q] = hole ? two : src[q];
            if (lane < r) vl.vol.tail[(row + x0 + j) * Rp + lane] = hole ? 2.f : ctile[j * DnP + Dm + lane];
        }
    }
    // ---- right view: diagonals of the tile.  Right pixel c gets d in [x0 - c, x0 + npix - 1 - c] (clipped to [0, Dn)),
    // a run of at most COST_TX consecutive d.  Main part (d < Dm): the run is at most COST_TX <= 64 long, i.e. exactly two
    // predicated warp stores per column starting at the run's first d (no loop, no alignment prologue: the stores are 4-byte
    // aligned 128-byte pieces either way).  Tail part (d >= Dm, usually the single last disparity): lanes over COLUMNS.
    {
        const int stride = DnP + 1;
        float* const rmain = vr.vol.main + row * Dm;
        static_assert(COST_TX <= 64, "two warp stores cover a run");
        for (int i = warp; i < COST_TX + Dn - 1; i += COST_WARPS) {
            const int c = cbase + i;
            if ((unsigned)c >= (unsigned)W) continue;
            const int d0 = max(0, x0 - c), d1 = min(Dm - 1, x0 + npix - 1 - c);
            const float* src = ctile + (c - x0) * DnP;  // + d * stride
            float* const pm = rmain + (size_t)c * Dm;
            const bool hole = MASK && vr.img4[row + c] == 0u;
            const int da = d0 + lane, db = da + 32;
            if (da <= d1) pm[da] = hole ? 2.f : src[da * stride];
            if (db <= d1) pm[db] = hole ? 2.f : src[db * stride];
        }
        if (r > 0) {
            float* const rtail = vr.vol.tail + row * Rp;
            for (int i = threadIdx.x; i < COST_TX + Dn - 1; i += COST_WARPS * 32) {
                const int c = cbase + i;
                if ((unsigned)c >= (unsigned)W) continue;
                const int d0 = max(Dm, x0 - c), d1 = min(Dn - 1, x0 + npix - 1 - c);
                const bool hole = MASK && vr.img4[row + c] == 0u;
                for (int d = d0; d <= d1; ++d) rtail[(size_t)c * Rp + d - Dm] = hole ? 2.f : ctile[(c - x0 + d) * DnP + d];
            }
        }
    }
    // ---- right view, cells whose left pixel c + d lies beyond the image: 2.f (ADCensus.cpp:562-566).  Last tile of the row.
    if (blockIdx.x == gridDim.x - 1) {
        for (int c = max(0, W - Dn + 1) + warp; c < W; c += COST_WARPS) {
            float* dst = vr.vol.main + (row + c) * Dm;
            float* dst_t = vr.vol.tail + (row + c) * Rp;
            for (int d = ((W - c) & ~31) + lane; d < Dn; d += 32) {
                if (d >= W - c) {
                    if (d < Dm) dst[d] = 2.f;
                    else dst_t[d - Dm] = 2.f;
                }
            }
        }
    }
}

template <int COST_TX, bool HSI, bool MASK>
static void launch_cost(const Launcher& L, const Dims& d, const ViewPtrs& left, const ViewPtrs& right, const float* d_tab_ad,
                        const float* d_tab_census)
{
    constexpr int tab_ad_n = HSI ? kTabAdHsi + 1 : kTabAdRgb + 2;
    constexpr int tab_c_pad = HSI ? TAB_C_PAD_HSI : TAB_C_PAD_RGB;
    const size_t ncol = (size_t)((COST_TX + d.Dn - 1 + 3) & ~3), dnp = (size_t)((d.Dn + 3) & ~3);
    const size_t smem = (size_t)(tab_ad_n + tab_c_pad) * 4 + 13 * ncol * 4 + (size_t)COST_TX * dnp * 4;
    static PerDevice smem_set;
    if (smem > smem_set.cur()) {
        cudaFuncSetAttribute(k_cost_init<COST_TX, HSI, MASK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        smem_set.cur() = smem;
    }
    dim3 grid((d.W + COST_TX - 1) / COST_TX, d.H);
    k_cost_init<COST_TX, HSI, MASK><<<grid, COST_WARPS * 32, smem, L.stream>>>(d, left, right, d_tab_ad, d_tab_census);
    L.count(1);
}

// ---- general form (slow path) ----------------------------------------------------------------------------
// costInitialize for minD != 0 (ADCensus.cpp:556-561): plane p of the left volume holds the cost of the pixel pair
// (xL, xR) = (x - minD, x - p), plane p of the right volume that of (x + p, x + minD) -- the matched pair is p - minD
// apart (negative for p < minD), not p, and the right volume is no longer a shear of the left one inside the image.
// One warp per pixel of one view, lanes over the planes; signatures come straight from global memory (L2-resident).
template <bool HSI, bool MASK>
__global__ void __launch_bounds__(256)
k_cost_init_general(Dims dm, ViewPtrs vl, ViewPtrs vr, const float* __restrict__ tab_ad, const float* __restrict__ tab_c)
{
    const int H = dm.H, W = dm.W, Dn = dm.Dn, m = dm.minD;
    const int lane = threadIdx.x & 31, x = blockIdx.x * 8 + (threadIdx.x >> 5), y = blockIdx.y, view = blockIdx.z;
    if (x >= W) return;
    const size_t npx = (size_t)H * W, row = (size_t)y * W;
    const int hw = kCensusW / 2;
    const bool yout = (y - kCensusH / 2 < 0) || (y + kCensusH / 2 >= H);
    constexpr int TAB_AD_USED = HSI ? kTabAdHsi : kTabAdRgb;
    const ViewPtrs& own = view ? vr : vl;
    const bool hole = MASK && own.img4[row + x] == 0u;  // a black pixel of the own view costs 2 (:551-555)
    auto load = [&](const ViewPtrs& v, int c, Sig& s) {
        const bool ok = !yout && c - hw >= 0 && c + hw < W;
        s.pix = ok ? v.img4[row + c] : kInvalidPix;
#pragma unroll
        for (int p = 0; p < 6; ++p) {
            const uint64_t w = ok ? v.census[(size_t)p * npx + row + c] : 0ull;
            s.w[2 * p] = (uint32_t)w;
            s.w[2 * p + 1] = (uint32_t)(w >> 32);
        }
    };
    // the pixel of the pair that does not depend on the plane: left view -> xL = x - minD, right view -> xR = x + minD
    Sig f;
    load(view ? vr : vl, view ? x + m : x - m, f);
    for (int p0 = 0; p0 < Dn; p0 += 32) {
        const int p = p0 + lane;
        if (p >= Dn) break;
        Sig g;
        load(view ? vl : vr, view ? x + p : x - p, g);  // left view: xR = x - p; right view: xL = x + p
        float cost = 2.f;
        if (!hole && f.pix != kInvalidPix && g.pix != kInvalidPix) {
            int ad3;
            if (HSI) {
                const uint32_t dv = __vabsdiffu4(f.pix, g.pix);
                const int hd = dv & 0xffu;
                ad3 = min(2 * min(hd, 255 - hd) + 5 * (int)(((dv >> 8) & 0xffu) + ((dv >> 16) & 0xffu)), TAB_AD_USED - 1);
            } else {
                ad3 = min((int)__vsadu4(f.pix, g.pix), TAB_AD_USED - 1);
            }
            float ec = tab_c[census_count(f, g)];
            if (MASK && (f.pix == 0u || g.pix == 0u)) ec = 0.f;
            cost = __fsub_rn(__fsub_rn(2.f, tab_ad[ad3]), ec);
        }
        *cell_ptr(own.vol, dm, row + x, p) = cost;
    }
}

void cost_init_general(const Launcher& L, const Dims& d, const ViewPtrs& left, const ViewPtrs& right, const float* d_tab_ad,
                       const float* d_tab_census, bool hsi, bool mask)
{
    dim3 grid((d.W + 7) / 8, d.H, 2);
    if (hsi) {
        if (mask) k_cost_init_general<true, true><<<grid, 256, 0, L.stream>>>(d, left, right, d_tab_ad, d_tab_census);
        else k_cost_init_general<true, false><<<grid, 256, 0, L.stream>>>(d, left, right, d_tab_ad, d_tab_census);
    } else {
        if (mask) k_cost_init_general<false, true><<<grid, 256, 0, L.stream>>>(d, left, right, d_tab_ad, d_tab_census);
        else k_cost_init_general<false, false><<<grid, 256, 0, L.stream>>>(d, left, right, d_tab_ad, d_tab_census);
    }
    L.count(1);
}

void cost_init(const Launcher& L, const Dims& d, const ViewPtrs& left, const ViewPtrs& right, const float* d_tab_ad,
               const float* d_tab_census, bool hsi, bool mask)
{
    const bool narrow = d.Dn > 256;
#define TSM_COST_CASE(TX, H, M) launch_cost<TX, H, M>(L, d, left, right, d_tab_ad, d_tab_census)
    if (mask) {
        if (hsi) { if (narrow) TSM_COST_CASE(32, true, true); else TSM_COST_CASE(64, true, true); }
        else { if (narrow) TSM_COST_CASE(32, false, true); else TSM_COST_CASE(64, false, true); }
    } else {
        if (hsi) { if (narrow) TSM_COST_CASE(32, true, false); else TSM_COST_CASE(64, true, false); }
        else { if (narrow) TSM_COST_CASE(32, false, false); else TSM_COST_CASE(64, false, false); }
    }
#undef TSM_COST_CASE
}

}  // namespace tsm
