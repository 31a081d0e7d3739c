// k_prep.cu -- per-view O(H*W) preparation kernels: BGRx packing, ternary census
// sign planes, cross arms, cross-window sizes and colour-similarity flags.
//
// Reference semantics (paths relative to the reference repo):
//   census sign test      source/ADCensus.cpp:461-472   (a4)
//   arm construction      source/ADCensus.cpp:604-659   (a6)
//   window sizes          source/ADCensus.cpp:716,733   (a7, d-independent)
//   P1/P2 similarity test source/ADCensus.cpp:927-934   (a8)
#include "tsm_common.cuh"

namespace tsm {

// Both views go through every preparation kernel in ONE launch (blockIdx.z or blockIdx.y = view): the stage is a chain of
// small latency-bound kernels, half as many launches with twice the CTAs each.
struct PrepView {
    const uint8_t* img;
    uint32_t* img4;
    uint64_t* census;
    uchar4* arms;
    uint32_t *desc_h, *desc_v, *fdesc_h, *fdesc_v;
    float *rcp_h, *rcp_v;
    uint8_t* flags;
};
struct PrepPair {
    PrepView v[2];
};

// ---- BGR -> BGRx ---------------------------------------------------------
__global__ void k_pack_bgrx(PrepPair pp, size_t npx)
{
    const uint8_t* __restrict__ img = pp.v[blockIdx.y].img;
    uint32_t* __restrict__ img4 = pp.v[blockIdx.y].img4;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npx) return;
    const uint8_t* p = img + i * 3;
    img4[i] = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16);
}

// ---- census sign planes ---------------------------------------------------
// One CTA = 32x8 pixel tile; the (8+6)x(32+8) BGRx window is staged in shared
// memory once and every thread builds its six 63-bit planes from it.
constexpr int CT_W = 32, CT_H = 8, CH_W = kCensusW / 2, CH_H = kCensusH / 2;
constexpr int CS_W = CT_W + 2 * CH_W, CS_H = CT_H + 2 * CH_H;

// HSI model (computeHSICensusCost, ADCensus.cpp:476-498): saturation and intensity keep the sign-product test;
// the hue term of a neighbour costs 0 only when BOTH views satisfy A = (dH <= -127 or 0 <= dH <= 127), i.e. it
// costs popc(nA_L | nA_R) with nA = not A.  Stored as plane "lt" = nA and plane "gt" = all ones, the generic
// match word (lt_L & gt_R) | (gt_L & lt_R) of k_cost_init evaluates exactly that.
template <bool HSI>
__global__ void __launch_bounds__(CT_W* CT_H)
k_census(PrepPair pp, int H, int W)
{
    const uint32_t* __restrict__ img4 = pp.v[blockIdx.z].img4;
    uint64_t* __restrict__ census = pp.v[blockIdx.z].census;
    __shared__ uint32_t tile[CS_H][CS_W + 1];
    const int x0 = blockIdx.x * CT_W, y0 = blockIdx.y * CT_H;
    const int tid = threadIdx.y * CT_W + threadIdx.x;
    for (int i = tid; i < CS_H * CS_W; i += CT_W * CT_H) {
        int ty = i / CS_W, tx = i % CS_W;
        int y = y0 + ty - CH_H, x = x0 + tx - CH_W;
        tile[ty][tx] = (y >= 0 && y < H && x >= 0 && x < W) ? img4[(size_t)y * W + x] : 0u;
    }
    __syncthreads();
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    if (x >= W || y >= H) return;
    uint64_t lt[3] = {0, 0, 0}, gt[3] = {0, 0, 0};
    if (y - CH_H >= 0 && y + CH_H < H && x - CH_W >= 0 && x + CH_W < W) {
        const uint32_t c = tile[threadIdx.y + CH_H][threadIdx.x + CH_W];
        const int cb = c & 0xff, cg = (c >> 8) & 0xff, cr = (c >> 16) & 0xff;
        int bit = 0;
#pragma unroll
        for (int i = 0; i < kCensusH; ++i)
#pragma unroll
            for (int j = 0; j < kCensusW; ++j, ++bit) {
                const uint32_t a = tile[threadIdx.y + i][threadIdx.x + j];
                const int ab = a & 0xff, ag = (a >> 8) & 0xff, ar = (a >> 16) & 0xff;
                const uint64_t m = 1ull << bit;
                if (HSI) {
                    const int dh = ab - cb;
                    if (!(dh <= -127 || (dh >= 0 && dh <= 127))) lt[0] |= m;
                    gt[0] |= m;
                } else {
                    if (ab < cb) lt[0] |= m;
                    if (ab > cb) gt[0] |= m;
                }
                if (ag < cg) lt[1] |= m;
                if (ag > cg) gt[1] |= m;
                if (ar < cr) lt[2] |= m;
                if (ar > cr) gt[2] |= m;
            }
    }
    const size_t npx = (size_t)H * W, p = (size_t)y * W + x;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        census[(size_t)c * npx + p] = lt[c];
        census[(size_t)(3 + c) * npx + p] = gt[c];
    }
}

// ---- cross arms -------------------------------------------------------------
__device__ __forceinline__ int arm_length(const uint32_t* __restrict__ img4, int H, int W, int y, int x, int dy, int dx,
                                          const ModelParams mp)
{
    // computeLimit, ADCensus.cpp:609-658, restated (the literal loop keeps a pixel counter d, an "inside" flag for the NEXT pixel and
    // three conditions; unrolled by hand it says):
    //   n_in  = pixels between p and the border in this direction;  0 -> arm 0
    //   ok_k  = cd(p, p_k) < tau1 && cd(p_k, p_{k-1}) < tau1 && k < L1 && (k <= L2 || cd(p, p_k) < tau2)
    //   the walk stops at the first k with  !ok_k  or  k + 1 > n_in (the next pixel is outside)  or  p_k black (mask matching, :624-629);
    //   arm = k - 1  (one pixel early at the border).
    // RGB: max-channel colour difference; HSI: the intensity tests (the hue and saturation assignments before them are overwritten,
    // :631-645).  A running pointer and the step bound n_in replace the 64-bit index and the four-way inside test of every step.
    const uint32_t* q = img4 + (size_t)y * W + x;
    const uint32_t p = *q;
    if (mp.mask && p == 0u) return 0;  // computeLimits, ADCensus.cpp:672-677
    const int n_in = dy < 0 ? y : dy > 0 ? H - 1 - y : dx < 0 ? x : W - 1 - x;
    if (n_in == 0) return 0;
    const ptrdiff_t step = (ptrdiff_t)dy * W + dx;
    uint32_t prev = p;
    if (!mp.hsi) {
        // RGB: "max over the channels of |a - b| < tau" == "no byte of vabsdiff4(a, b) exceeds tau - 1" (byte 3 of both words is 0);
        // any-byte-greater-than-n for n <= 127 is ((x + 0x01010101 * (127 - n)) | x) & 0x80808080 -- a carry out of a byte means
        // that byte already exceeds n, so it cannot turn a "no" into a "yes".  Three instructions per test instead of eight on the
        // ALU pipe this kernel is bound by (math-pipe throttle was its top stall).
        const uint32_t c1 = 0x01010101u * (uint32_t)(128 - mp.tau1), c2 = 0x01010101u * (uint32_t)(128 - mp.tau2);
        for (int k = 1;; ++k) {
            q += step;
            const uint32_t p1 = *q;
            const uint32_t ad = __vabsdiffu4(p, p1), ad2 = __vabsdiffu4(p1, prev);
            const bool lt1 = (((ad + c1) | ad) & 0x80808080u) == 0u;    // cd  < tau1
            const bool lt1b = (((ad2 + c1) | ad2) & 0x80808080u) == 0u; // cd2 < tau1
            const bool lt2 = (((ad + c2) | ad) & 0x80808080u) == 0u;    // cd  < tau2
            const bool ok = lt1 && lt1b && k < mp.L1 && (k <= mp.L2 || lt2);
            if (!ok || k + 1 > n_in || (mp.mask && p1 == 0u)) return k - 1;
            prev = p1;
        }
    }
    for (int k = 1;; ++k) {
        q += step;
        const uint32_t p1 = *q;
        const int cd = abs((int)((p >> 16) & 0xffu) - (int)((p1 >> 16) & 0xffu));
        const int cd2 = abs((int)((p1 >> 16) & 0xffu) - (int)((prev >> 16) & 0xffu));
        const bool ok = cd < mp.tau1 && cd2 < mp.tau1 && k < mp.L1 && (k <= mp.L2 || cd < mp.tau2);
        if (!ok || k + 1 > n_in || (mp.mask && p1 == 0u)) return k - 1;
        prev = p1;
    }
}

// One thread = one pixel and one AXIS (blockIdx.z = 2 * view + axis): the vertical thread walks up and down and also writes
// the pixel's similarity flags, the horizontal one walks left and right.  (One thread for all four walks measured 0.25 ms per
// view at 1080p: the walks are serial chains of dependent L1 loads and a warp waits for its longest arm four times over.)
__global__ void k_arms_flags(PrepPair pp, int H, int W, ModelParams mp)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    const PrepView& v = pp.v[blockIdx.z >> 1];
    const uint32_t* __restrict__ img4 = v.img4;
    const size_t p = (size_t)y * W + x;
    unsigned char* a = reinterpret_cast<unsigned char*>(v.arms + p);  // (up, down, left, right)
    if ((blockIdx.z & 1) == 0) {
        uchar2 ud;
        ud.x = (unsigned char)arm_length(img4, H, W, y, x, -1, 0, mp);
        ud.y = (unsigned char)arm_length(img4, H, W, y, x, 1, 0, mp);
        *reinterpret_cast<uchar2*>(a) = ud;
        // similarity flags of computeP1P2 (:927-934)
        const uint32_t c = img4[p];
        uint8_t f = 0;
        if (y > 0 && (mp.hsi ? hue_diff_u32(c, img4[p - W]) : color_diff_u32(c, img4[p - W])) < mp.sim) f |= 1;
        if (x > 0 && (mp.hsi ? hue_diff_u32(c, img4[p - 1]) : color_diff_u32(c, img4[p - 1])) < mp.sim) f |= 2;
        if (mp.mask && c == 0u) f |= 4;  // mask matching: a black pixel (ADCensus.cpp:824, 862)
        v.flags[p] = f;
    } else {
        uchar2 lr;
        lr.x = (unsigned char)arm_length(img4, H, W, y, x, 0, -1, mp);
        lr.y = (unsigned char)arm_length(img4, H, W, y, x, 0, 1, mp);
        *reinterpret_cast<uchar2*>(a + 2) = lr;
    }
}

// ---- aggregation step descriptors -----------------------------------------------
// aggregation1D propagates the window size like a cost (ADCensus.cpp:716) starting from all
// ones (:733): horizontal-first iteration -> N_hf = sum over the vertical arm of the row
// lengths; vertical-first -> N_vf = sum over the horizontal arm of the column lengths.
// The pass that ENDS an iteration divides by N (:743-749): a vertical pass ends a
// horizontal-first iteration (N_hf), a horizontal pass a vertical-first one (N_vf).
__global__ void k_agg_desc(PrepPair pp, int H, int W, int Wd, int Hd)
{
    const PrepView& v = pp.v[blockIdx.z];
    const uchar4* __restrict__ arms = v.arms;
    uint32_t* __restrict__ desc_h = v.desc_h;
    uint32_t* __restrict__ desc_v = v.desc_v;
    float* __restrict__ rcp_h = v.rcp_h;
    float* __restrict__ rcp_v = v.rcp_v;
    uint32_t* __restrict__ fdesc_h = v.fdesc_h;
    uint32_t* __restrict__ fdesc_v = v.fdesc_v;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    const size_t p = (size_t)y * W + x;
    const uchar4 a = arms[p];
    int nh = 0, nv = 0;
    for (int j = -(int)a.x; j <= (int)a.y; ++j) {
        const uchar4 q = arms[p + (ptrdiff_t)j * W];
        nh += q.z + q.w + 1;
    }
    for (int j = -(int)a.z; j <= (int)a.w; ++j) {
        const uchar4 q = arms[p + j];
        nv += q.x + q.y + 1;
    }
    desc_h[(size_t)y * Wd + x] = (uint32_t)a.z | ((uint32_t)a.w << 8) | ((uint32_t)nv << 16);
    desc_v[(size_t)x * Hd + y] = (uint32_t)a.x | ((uint32_t)a.y << 8) | ((uint32_t)nh << 16);
    // fused walk (k_agg_fused): distances from the newest prefix to P[o + b + 1] and P[o - a] in ring units of 3
    fdesc_h[(size_t)y * Wd + x] = 3u * (uint32_t)(kMaxArm - a.w) | (3u * (uint32_t)(kMaxArm + 1 + a.z)) << 8 | ((uint32_t)nv << 16);
    fdesc_v[(size_t)x * Hd + y] = 3u * (uint32_t)(kMaxArm - a.y) | (3u * (uint32_t)(kMaxArm + 1 + a.x)) << 8 | ((uint32_t)nh << 16);
    // correctly rounded reciprocals of the divisors: the normalising passes divide with one residual correction
    rcp_h[(size_t)y * Wd + x] = __frcp_rn((float)nv);
    rcp_v[(size_t)x * Hd + y] = __frcp_rn((float)nh);
}

// ---- scan tables for the scanline kernels -------------------------------------------
// stab[plane][y][kTfPad + c], c in [-kTfPad, pitch - kTfPad):
//   bits 0..23: flag bit `plane` of the OTHER image at (y, c + s*(32k + minD)), k = 0..K-1, 0 outside the image --
//               a lane of the scanline warp that handles d = lane + 32k gets all its K similarity
//               bits from the one word at column x + s*lane;
//   bit 31    : flag bit `plane` of the OWN image at (y, c);
//   bits 30/29: mask matching only, see below.
__global__ void k_scan_table(const uint8_t* __restrict__ flags_left, const uint8_t* __restrict__ flags_right, uint32_t* __restrict__ stab_left,
                             uint32_t* __restrict__ stab_right, int H, int W, int Wp, int K, int minD)
{
    // blockIdx.z = view: the left volume looks at the right image at x + d, the right volume at the left image at x - d
    const uint8_t* __restrict__ fown = blockIdx.z ? flags_right : flags_left;
    const uint8_t* __restrict__ foth = blockIdx.z ? flags_left : flags_right;
    uint32_t* __restrict__ stab = blockIdx.z ? stab_right : stab_left;
    const int s = blockIdx.z ? -1 : 1;
    const int cx = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (cx >= Wp) return;
    const int c = cx - kTfPad;
    unsigned tv = 0, th = 0;
    for (int k = 0; k < K; ++k) {
        const int x = c + s * (32 * k + minD);  // plane p = lane + 32 k looks at x +- (p + minD), ADCensus.cpp:890
        if (x >= 0 && x < W) {
            const unsigned f = foth[(size_t)y * W + x];
            tv |= (f & 1u) << k;
            th |= ((f >> 1) & 1u) << k;
        }
    }
    if (c >= 0 && c < W) {
        const unsigned f = fown[(size_t)y * W + c];
        tv |= (f & 1u) << 31;
        th |= ((f >> 1) & 1u) << 31;
        // mask matching: a step whose PREDECESSOR is black is skipped.  (y, c) is the flag pixel of the step = the
        // later of (pixel, predecessor): bit 30 = black predecessor of a forward step, bit 29 of a backward step.
        const unsigned own_black = (f >> 2) & 1u;
        tv |= own_black << 29;
        th |= own_black << 29;
        if (y > 0) tv |= ((fown[(size_t)(y - 1) * W + c] >> 2) & 1u) << 30;
        if (c > 0) th |= ((fown[(size_t)y * W + c - 1] >> 2) & 1u) << 30;
    }
    stab[(size_t)y * Wp + cx] = tv;
    stab[(size_t)H * Wp + (size_t)y * Wp + cx] = th;
}

void prep_scan_tables(const Launcher& L, const Dims& d, const uint8_t* flags_left, const uint8_t* flags_right,
                      uint32_t* stab_left, uint32_t* stab_right)
{
    const int Wp = d.stab_pitch(), K = (d.Dn + 31) / 32;
    dim3 tg((Wp + 127) / 128, d.H, 2);
    k_scan_table<<<tg, 128, 0, L.stream>>>(flags_left, flags_right, stab_left, stab_right, d.H, d.W, Wp, K, d.minD);
    L.count(1);
}

// ---- HSI preprocessing (ADCensus::compute, ADCensus.cpp:350-371) ---------------------------------------
// bgr2hsi (:1429-1473) is a pure function of the 24-bit pixel with sqrtf / acosf inside; the host evaluates it
// for all 2^24 inputs once per process with its own libm (tsm_capi.cu) -- the same trick as the exp() tables --
// so the conversion is a table lookup and bit-identical to the reference.
// `filter` (ROI mode, bgr2hsi(..., true), :1463-1470): pixels whose hue is >= 60 or <= 10 become (0, 0, 0).
__global__ void k_hsi_lookup(PrepPair pp, int to_arms, const uint32_t* __restrict__ lut, size_t n, int filter)
{
    // in: the view's BGRx image; out: the same buffer (ROI mode) or the arms buffer as scratch (free until k_arms_flags, same size)
    const uint32_t* __restrict__ bgr4 = pp.v[blockIdx.y].img4;
    uint32_t* __restrict__ hsi4 = to_arms ? reinterpret_cast<uint32_t*>(pp.v[blockIdx.y].arms) : pp.v[blockIdx.y].img4;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t v = lut[bgr4[i] & 0xffffffu];
    const uint32_t h = v & 0xffu;
    if (filter && (h >= 60u || h <= 10u)) v = 0u;
    hsi4[i] = v;
}

// computeGaussMedian(src, dst, 3) (:1475-1499): filter2D with the 3x3 kernel [1 2 1; 2 4 2; 1 2 1] / 16
// (cv::getGaussianKernel(3, -1)), BORDER_CONSTANT 0, rounded half to even; a channel takes the filtered value
// when it is far from it (hue: circular distance >= 2; S, I: >= 3).
__global__ void k_gauss_median(PrepPair pp, int H, int W)
{
    const uint32_t* __restrict__ src = reinterpret_cast<const uint32_t*>(pp.v[blockIdx.z].arms);  // scratch filled by k_hsi_lookup
    uint32_t* __restrict__ dst = pp.v[blockIdx.z].img4;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    int sum[3] = {0, 0, 0};
#pragma unroll
    for (int i = -1; i <= 1; ++i)
#pragma unroll
        for (int j = -1; j <= 1; ++j) {
            const int yy = y + i, xx = x + j;
            if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
            const uint32_t v = src[(size_t)yy * W + xx];
            const int w = (i == 0 ? 2 : 1) * (j == 0 ? 2 : 1);
            sum[0] += w * (int)(v & 0xffu);
            sum[1] += w * (int)((v >> 8) & 0xffu);
            sum[2] += w * (int)((v >> 16) & 0xffu);
        }
    const uint32_t s = src[(size_t)y * W + x];
    uint32_t out = 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int q = sum[c] >> 4, r = sum[c] & 15;
        const int med = q + ((r > 8 || (r == 8 && (q & 1))) ? 1 : 0);  // round(sum / 16) half to even, <= 255
        const int own = (int)((s >> (8 * c)) & 0xffu);
        int diff = abs(own - med);
        bool take;
        if (c == 0) { diff = min(diff, 255 - diff); take = diff >= 2; }
        else take = diff >= 3;
        out |= (uint32_t)(take ? med : own) << (8 * c);
    }
    dst[(size_t)y * W + x] = out;
}

void prep_views(const Launcher& L, const Dims& d, const uint8_t* const img[2], uint32_t* const img4[2], uint64_t* const census[2],
                uchar4* const arms[2], uint32_t* const desc_h[2], uint32_t* const desc_v[2], uint32_t* const fdesc_h[2],
                uint32_t* const fdesc_v[2], uint8_t* const flags[2], const ModelParams& mp, const uint32_t* hsi_lut, bool roi)
{
    const size_t npx = d.npx();
    PrepPair pp;
    for (int k = 0; k < 2; ++k) {
        PrepView& v = pp.v[k];
        v.img = img[k]; v.img4 = img4[k]; v.census = census[k]; v.arms = arms[k];
        v.desc_h = desc_h[k]; v.desc_v = desc_v[k];
        v.rcp_h = reinterpret_cast<float*>(desc_h[k] + d.desc_h_words());
        v.rcp_v = reinterpret_cast<float*>(desc_v[k] + d.desc_v_words());
        v.fdesc_h = fdesc_h[k] + kFdescFront; v.fdesc_v = fdesc_v[k] + kFdescFront;
        v.flags = flags[k];
    }
    dim3 b(32, 8), g((d.W + 31) / 32, (d.H + 7) / 8, 2), g1((unsigned)((npx + 255) / 256), 2);
    k_pack_bgrx<<<g1, 256, 0, L.stream>>>(pp, npx);
    if (mp.hsi && roi) {  // ADCensus.cpp:354-360: hue-filtered HSI, no Gauss-median
        k_hsi_lookup<<<g1, 256, 0, L.stream>>>(pp, 0, hsi_lut, npx, 1);
        L.count(1);
    } else if (mp.hsi) {  // :361-370
        k_hsi_lookup<<<g1, 256, 0, L.stream>>>(pp, 1, hsi_lut, npx, 0);
        k_gauss_median<<<g, b, 0, L.stream>>>(pp, d.H, d.W);
        L.count(2);
    }
    dim3 cb(CT_W, CT_H), cg((d.W + CT_W - 1) / CT_W, (d.H + CT_H - 1) / CT_H, 2);
    if (mp.hsi) k_census<true><<<cg, cb, 0, L.stream>>>(pp, d.H, d.W);
    else k_census<false><<<cg, cb, 0, L.stream>>>(pp, d.H, d.W);
    dim3 ga(g.x, g.y, 4);  // 2 views x 2 axes
    k_arms_flags<<<ga, b, 0, L.stream>>>(pp, d.H, d.W, mp);
    k_agg_desc<<<g, b, 0, L.stream>>>(pp, d.H, d.W, d.Wd(), d.Hd());
    L.count(4);
}

__global__ void k_roi_finish(float* __restrict__ fin, const uint8_t* __restrict__ bgr, size_t n, float offset)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float d = fin[i];
    if (d > 0.f) d = __fadd_rn(d, offset);  // disparityOffset: d + offset (int -> float), ADCensus.cpp:1422-1423
    const bool black = bgr[3 * i] == 0 && bgr[3 * i + 1] == 0 && bgr[3 * i + 2] == 0;
    if ((black && d > 0.f) || d == 0.f) d = -1.f;  // :397-400
    fin[i] = d;
}

void roi_finish(const Launcher& L, const Dims& d, float* fin, const uint8_t* left_bgr, int offset)
{
    const size_t n = d.npx();
    k_roi_finish<<<(unsigned)((n + 255) / 256), 256, 0, L.stream>>>(fin, left_bgr, n, (float)offset);
    L.count(1);
}

}  // namespace tsm
