/*
 * oracle/adcensus_oracle.c -- TEST INFRASTRUCTURE (oracle), not product code.
 * See adcensus_oracle.h.  Every function cites the reference lines it restates
 * (paths relative to /root/reference).  Build: oracle/Makefile (`make port`),
 * flags -O2 -fopenmp -ffp-contract=off (no FMA contraction, no fast-math).
 */
#include "adcensus_oracle.h"
#include "cvport.h"
#include <float.h>
#include <math.h>
#include <omp.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

static inline int iabs(int v) { return v < 0 ? -v : v; }
static inline int imax(int a, int b) { return a > b ? a : b; }

/* colorDiff, RGB branch: max over channels of |dc|.  source/ADCensus.cpp:583-602 */
static inline int color_diff(const uint8_t* p, const uint8_t* q)
{
    int d = iabs((int)p[0] - (int)q[0]);
    d = imax(d, iabs((int)p[1] - (int)q[1]));
    d = imax(d, iabs((int)p[2] - (int)q[2]));
    return d;
}

/* ------------------------------------------------------------------ a3 / a4 */

/* source/ADCensus.cpp:426-437 (integer part; the reference then divides by 3.f) */
int orc_ad3(const uint8_t* L, const uint8_t* R, int W, int y, int xl, int xr)
{
    const uint8_t* p = L + ((size_t)y * W + xl) * 3;
    const uint8_t* q = R + ((size_t)y * W + xr) * 3;
    return iabs(p[0] - q[0]) + iabs(p[1] - q[1]) + iabs(p[2] - q[2]);
}

/* source/ADCensus.cpp:454-474, literally: count of (Lnb-Lc)*(Rnb-Rc) < 0. */
int orc_census_direct(const uint8_t* L, const uint8_t* R, int W, int y, int xl, int xr)
{
    const uint8_t* lc = L + ((size_t)y * W + xl) * 3;
    const uint8_t* rc = R + ((size_t)y * W + xr) * 3;
    int n = 0;
    for (int i = -ORC_CENSUS_H / 2; i <= ORC_CENSUS_H / 2; ++i)
        for (int j = -ORC_CENSUS_W / 2; j <= ORC_CENSUS_W / 2; ++j) {
            const uint8_t* la = L + ((size_t)(y + i) * W + xl + j) * 3;
            const uint8_t* ra = R + ((size_t)(y + i) * W + xr + j) * 3;
            for (int k = 0; k < 3; ++k)
                n += (((int)la[k] - (int)lc[k]) * ((int)ra[k] - (int)rc[k]) < 0) ? 1 : 0;
        }
    return n;
}

/* Sign planes equivalent to ADCensus.cpp:461-472: the product of two differences is
 * negative iff one is <0 and the other >0, so
 *   census = sum_c popcount((ltL_c & gtR_c) | (gtL_c & ltR_c)). */
void orc_census_signatures(const uint8_t* img, int H, int W, uint64_t* lt, uint64_t* gt)
{
    const int hh = ORC_CENSUS_H / 2, hw = ORC_CENSUS_W / 2;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            uint64_t l[3] = {0, 0, 0}, g[3] = {0, 0, 0};
            if (y - hh >= 0 && y + hh < H && x - hw >= 0 && x + hw < W) {
                const uint8_t* c = img + ((size_t)y * W + x) * 3;
                int bit = 0;
                for (int i = -hh; i <= hh; ++i)
                    for (int j = -hw; j <= hw; ++j, ++bit) {
                        const uint8_t* a = img + ((size_t)(y + i) * W + x + j) * 3;
                        for (int k = 0; k < 3; ++k) {
                            if (a[k] < c[k]) l[k] |= 1ull << bit;
                            if (a[k] > c[k]) g[k] |= 1ull << bit;
                        }
                    }
            }
            for (int k = 0; k < 3; ++k) {
                lt[((size_t)y * W + x) * 3 + k] = l[k];
                gt[((size_t)y * W + x) * 3 + k] = g[k];
            }
        }
}

/* ------------------------------------------------------------------------ a5 */

/* source/ADCensus.cpp:500-581.  cost = (2.f - expf(-ad/10.f)) - expf(-census/30.f)
 * (ADCensus.cpp:518), ad = (float)sum/3.f (:435).  ad has 766 and census 187 distinct
 * inputs, so both exponentials are tabulated with the host's expf.  Border rule :562-566. */
void orc_cost_init(const uint8_t* L, const uint8_t* R, int H, int W, int Dn, float* vol0, float* vol1)
{
    float tab_ad[766], tab_c[190];
    for (int s = 0; s < 766; ++s) {
        float ad = (float)s / 3.f;
        tab_ad[s] = expf(-ad / ORC_LAMBDA_AD);
    }
    for (int n = 0; n < 190; ++n) tab_c[n] = expf(-(float)n / ORC_LAMBDA_CENSUS);
    uint64_t* ltL = (uint64_t*)malloc((size_t)H * W * 3 * 8);
    uint64_t* gtL = (uint64_t*)malloc((size_t)H * W * 3 * 8);
    uint64_t* ltR = (uint64_t*)malloc((size_t)H * W * 3 * 8);
    uint64_t* gtR = (uint64_t*)malloc((size_t)H * W * 3 * 8);
    orc_census_signatures(L, H, W, ltL, gtL);
    orc_census_signatures(R, H, W, ltR, gtR);
    const int hh = ORC_CENSUS_H / 2, hw = ORC_CENSUS_W / 2;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; ++y) {
        const int yout = (y - hh < 0) || (y + hh >= H);
        for (int x = 0; x < W; ++x)
            for (int view = 0; view < 2; ++view) {
                float* out = (view ? vol1 : vol0) + ((size_t)y * W + x) * Dn;
                for (int d = 0; d < Dn; ++d) {
                    /* view 0: (colL, colR) = (x, x-d); view 1: (x+d, x)   (:556-561) */
                    const int xl = view ? x + d : x, xr = view ? x : x - d;
                    const int oob = yout || xl - hw < 0 || xl + hw >= W || xr - hw < 0 || xr + hw >= W;
                    if (oob) { out[d] = 2.f; continue; }
                    const size_t il = ((size_t)y * W + xl) * 3, ir = ((size_t)y * W + xr) * 3;
                    int ad3 = iabs(L[il] - R[ir]) + iabs(L[il + 1] - R[ir + 1]) + iabs(L[il + 2] - R[ir + 2]);
                    int cen = 0;
                    for (int k = 0; k < 3; ++k)
                        cen += __builtin_popcountll((ltL[il + k] & gtR[ir + k]) | (gtL[il + k] & ltR[ir + k]));
                    out[d] = (2.f - tab_ad[ad3]) - tab_c[cen];
                }
            }
    }
    free(ltL); free(gtL); free(ltR); free(gtR);
}

/* ------------------------------------------------------------------------ a6 */

/* source/ADCensus.cpp:604-659 (computeLimit), RGB branch, walked literally. */
static int arm_length(const uint8_t* img, int H, int W, int y, int x, int dy, int dx)
{
    const uint8_t* p = img + ((size_t)y * W + x) * 3;
    int d = 1;
    int y1 = y + dy, x1 = x + dx;
    const uint8_t* p2 = p;
    int inside = (0 <= y1) && (y1 < H) && (0 <= x1) && (x1 < W);
    if (inside) {
        int color_cond = 1, wlimit_cond = 1, fcolor_cond = 1;
        while (color_cond && wlimit_cond && fcolor_cond && inside) {
            const uint8_t* p1 = img + ((size_t)y1 * W + x1) * 3;
            color_cond = color_diff(p, p1) < ORC_TAU1 && color_diff(p1, p2) < ORC_TAU1;
            wlimit_cond = d < ORC_L1;
            fcolor_cond = (d <= ORC_L2) || (d > ORC_L2 && color_diff(p, p1) < ORC_TAU2);
            p2 = p1;
            y1 += dy;
            x1 += dx;
            inside = (0 <= y1) && (y1 < H) && (0 <= x1) && (x1 < W);
            d++;
        }
        d--;
    }
    return d - 1;
}

/* source/ADCensus.cpp:661-683 + 760-766 */
void orc_arms(const uint8_t* img, int H, int W, int32_t* up, int32_t* down, int32_t* left, int32_t* right)
{
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            size_t i = (size_t)y * W + x;
            up[i] = arm_length(img, H, W, y, x, -1, 0);
            down[i] = arm_length(img, H, W, y, x, 1, 0);
            left[i] = arm_length(img, H, W, y, x, 0, -1);
            right[i] = arm_length(img, H, W, y, x, 0, 1);
        }
}

/* ------------------------------------------------------------------------ a7 */

/* One aggregation1D pass (source/ADCensus.cpp:685-723) over a d-innermost volume:
 * out(p) = sum_{j=-a(p)}^{b(p)} in(p + j*dir), fp32, accumulator starts at 0.f and
 * adds in ascending j (:711-716).  Window sizes are propagated the same way. */
static void agg_pass(const float* in, float* out, const int32_t* ws_in, int32_t* ws_out, int H, int W, int Dn,
                     const int32_t* arm_neg, const int32_t* arm_pos, int dy, int dx)
{
#pragma omp parallel
    {
        float* acc = (float*)malloc((size_t)Dn * sizeof(float));
#pragma omp for schedule(static)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                const size_t p = (size_t)y * W + x;
                const int dmin = -arm_neg[p], dmax = arm_pos[p];
                for (int d = 0; d < Dn; ++d) acc[d] = 0.f;
                int ws = 0;
                for (int j = dmin; j <= dmax; ++j) {
                    const size_t q = (size_t)(y + j * dy) * W + (x + j * dx);
                    const float* src = in + q * Dn;
                    for (int d = 0; d < Dn; ++d) acc[d] += src[d];
                    ws += ws_in[q];
                }
                memcpy(out + p * Dn, acc, (size_t)Dn * sizeof(float));
                ws_out[p] = ws;
            }
        free(acc);
    }
}

/* aggregation2D x iterations (source/ADCensus.cpp:725-751, 768-784): iteration i
 * runs horizontal-first when i is even; after both passes divide by the window size. */
void orc_aggregate(float* vol, int H, int W, int Dn, const int32_t* up, const int32_t* down,
                   const int32_t* left, const int32_t* right)
{
    const size_t npx = (size_t)H * W;
    float* tmp = (float*)malloc(npx * Dn * sizeof(float));
    int32_t* ws0 = (int32_t*)malloc(npx * sizeof(int32_t));
    int32_t* ws1 = (int32_t*)malloc(npx * sizeof(int32_t));
    int hf = 1;
    for (int it = 0; it < ORC_ITERATIONS; ++it) {
        for (size_t i = 0; i < npx; ++i) ws0[i] = 1;
        if (hf) {
            agg_pass(vol, tmp, ws0, ws1, H, W, Dn, left, right, 0, 1);
            agg_pass(tmp, vol, ws1, ws0, H, W, Dn, up, down, 1, 0);
        } else {
            agg_pass(vol, tmp, ws0, ws1, H, W, Dn, up, down, 1, 0);
            agg_pass(tmp, vol, ws1, ws0, H, W, Dn, left, right, 0, 1);
        }
#pragma omp parallel for schedule(static)
        for (size_t i = 0; i < npx; ++i) {
            const float n = (float)ws0[i]; /* float /= int, :747 */
            float* c = vol + i * Dn;
            for (int d = 0; d < Dn; ++d) c[d] = c[d] / n;
        }
        hf = !hf;
    }
    free(tmp); free(ws0); free(ws1);
}

/* ------------------------------------------------------------------------ a8 */

/* partialOptimization + computeP1P2 for one pixel p with predecessor q
 * (source/ADCensus.cpp:869-981).  own_sim = colorDiff(own(p),own(q)) < 15.
 * oth_sim[x'] = colorDiff(oth(p row, x'), oth(q row, x'')) < 15 indexed by the p-side
 * column x' = xp + sgn*d; it is only consulted when lo <= x' <= hi (both columns
 * inside the image, :929-930), otherwise d2 = colorDiff+1 (:928). */
static void scan_pixel(float* cp, const float* cq, int Dn, int own_sim, const uint8_t* oth_sim, int xp, int sgn,
                       int lo, int hi, const float* P1, const float* P2)
{
    float m = cq[0];
    for (int d = 1; d < Dn; ++d)
        if (m > cq[d]) m = cq[d];
    if (m == 0) return; /* :880 */
    for (int d = 0; d < Dn; ++d) {
        const float cost = cp[d] - m;
        const int xo = xp + sgn * d;
        const int s = own_sim + ((xo >= lo && xo <= hi) ? oth_sim[xo] : 0);
        const float p1 = P1[s], p2 = P2[s];
        float mo = m + p2;
        float t = cq[d];
        if (mo > t) mo = t;
        if (d != 0) {
            t = cq[d - 1] + p1;
            if (mo > t) mo = t;
        }
        if (d != Dn - 1) {
            t = cq[d + 1] + p1;
            if (mo > t) mo = t;
        }
        cp[d] = (cost + mo) / 2;
    }
}

/* scanline (source/ADCensus.cpp:983-995): down, up, right(w=1..W-1), left(w=W-2..0),
 * each pass in place and strictly sequential along its path (the reference's OpenMP
 * pragmas at :801-815 / :837-853 race; parity = sequential semantics). */
void orc_scanline(float* vol, int H, int W, int Dn, const uint8_t* own, const uint8_t* other, int view)
{
    /* P1/P2 classes, computeP1P2 :954-979; index = number of similar image pairs */
    const float P1[3] = {ORC_PI1 / 10.f, ORC_PI1 / 4.f, ORC_PI1};
    const float P2[3] = {ORC_PI2 / 10.f, ORC_PI2 / 4.f, ORC_PI2};
    const int sgn = view == 0 ? 1 : -1; /* :919-924 disparity sign */
    const size_t npx = (size_t)H * W;
    /* fv[y][x] = cd(I(y,x), I(y-1,x)) < 15 ; fh[y][x] = cd(I(y,x), I(y,x-1)) < 15 */
    uint8_t* fv_own = (uint8_t*)calloc(npx, 1);
    uint8_t* fv_oth = (uint8_t*)calloc(npx, 1);
    uint8_t* fh_own = (uint8_t*)calloc(npx, 1);
    uint8_t* fh_oth = (uint8_t*)calloc(npx, 1);
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            const size_t i = (size_t)y * W + x;
            if (y > 0) {
                fv_own[i] = color_diff(own + i * 3, own + (i - W) * 3) < ORC_COLOR_DIFF;
                fv_oth[i] = color_diff(other + i * 3, other + (i - W) * 3) < ORC_COLOR_DIFF;
            }
            if (x > 0) {
                fh_own[i] = color_diff(own + i * 3, own + (i - 1) * 3) < ORC_COLOR_DIFF;
                fh_oth[i] = color_diff(other + i * 3, other + (i - 1) * 3) < ORC_COLOR_DIFF;
            }
        }
    /* vertical, top -> bottom (:987) then bottom -> top (:989); columns are independent */
#pragma omp parallel
    {
        for (int y = 1; y < H; ++y) {
#pragma omp for schedule(static)
            for (int x = 0; x < W; ++x) {
                const size_t p = (size_t)y * W + x, q = p - W;
                scan_pixel(vol + p * Dn, vol + q * Dn, Dn, fv_own[p], fv_oth + (size_t)y * W, x, sgn, 0, W - 1, P1, P2);
            }
        }
        for (int y = H - 2; y >= 0; --y) {
#pragma omp for schedule(static)
            for (int x = 0; x < W; ++x) {
                const size_t p = (size_t)y * W + x, q = p + W;
                scan_pixel(vol + p * Dn, vol + q * Dn, Dn, fv_own[q], fv_oth + (size_t)(y + 1) * W, x, sgn, 0, W - 1, P1, P2);
            }
        }
        /* horizontal (:991, :993); rows are independent */
#pragma omp for schedule(static)
        for (int y = 0; y < H; ++y) {
            for (int x = 1; x < W; ++x) {
                const size_t p = (size_t)y * W + x, q = p - 1;
                /* columns x+sgn*d and x-1+sgn*d inside  <=>  1 <= x+sgn*d <= W-1 */
                scan_pixel(vol + p * Dn, vol + q * Dn, Dn, fh_own[p], fh_oth + (size_t)y * W, x, sgn, 1, W - 1, P1, P2);
            }
            for (int x = W - 2; x >= 0; --x) {
                const size_t p = (size_t)y * W + x, q = p + 1;
                /* columns x+sgn*d and x+1+sgn*d inside; flag lives at the q-side column */
                scan_pixel(vol + p * Dn, vol + q * Dn, Dn, fh_own[q], fh_oth + (size_t)y * W + 1, x, sgn, 0, W - 2, P1, P2);
            }
        }
    }
    free(fv_own); free(fv_oth); free(fh_own); free(fh_oth);
}

/* ------------------------------------------------------------------ a9 / a10 */

/* source/ADCensus.cpp:1394-1413 with minD = 0: first strict minimum over d = 0..maxD. */
void orc_wta(const float* vol, int H, int W, int Dn, int32_t* disp)
{
#pragma omp parallel for schedule(static)
    for (size_t p = 0; p < (size_t)H * W; ++p) {
        const float* c = vol + p * Dn;
        float low = FLT_MAX;
        int best = 0; /* reference leaves disp uninitialised if nothing < FLT_MAX; cannot happen */
        for (int d = 0; d < Dn; ++d)
            if (low > c[d]) { low = c[d]; best = d; }
        disp[p] = best;
    }
}

/* source/ADCensus.cpp:1013-1044, dispTolerance = 0, minD = 0 */
void orc_lrc(const int32_t* dl, const int32_t* dr, int H, int W, int maxD, int32_t* out)
{
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            int disp = dl[(size_t)y * W + x];
            if (x - disp < 0 || iabs(disp - dr[(size_t)y * W + x - disp]) > ORC_DISP_TOL) {
                int occlusion = 1;
                for (int d = 0; d <= maxD; ++d)
                    if (x - d >= 0 && d == dr[(size_t)y * W + x - d]) { occlusion = 0; break; }
                disp = occlusion ? ORC_OCCLUSION : ORC_MISMATCH;
            }
            out[(size_t)y * W + x] = disp;
        }
}

/* ----------------------------------------------------------------------- a11 */

/* source/ADCensus.cpp:1046-1159.  hist is cleared only in the high-vote branch
 * (:1150), so votes of low-vote outliers leak into the next high-vote outlier in
 * raster order.  Reproduced literally (serial). */
void orc_region_voting(int32_t* disp, int H, int W, int Dn, const int32_t* up, const int32_t* down,
                       const int32_t* left, const int32_t* right, int horizontal_first)
{
    const size_t npx = (size_t)H * W;
    int32_t* tmp = (int32_t*)malloc(npx * sizeof(int32_t));
    int* hist = (int*)calloc((size_t)Dn, sizeof(int));
    const int32_t *outerA, *outerB, *innerA, *innerB;
    if (horizontal_first) { outerA = up; outerB = down; innerA = left; innerB = right; }
    else { outerA = left; outerB = right; innerA = up; innerB = down; }
    for (int h = 0; h < H; ++h)
        for (int w = 0; w < W; ++w) {
            const size_t p = (size_t)h * W + w;
            if (disp[p] >= 0) { tmp[p] = disp[p]; continue; }
            const int oa = -outerA[p], ob = outerB[p];
            int vote = 0;
            for (int outer = oa; outer <= ob; ++outer) {
                const size_t c = horizontal_first ? (size_t)(h + outer) * W + w : (size_t)h * W + (w + outer);
                const int ia = -innerA[c], ib = innerB[c];
                for (int inner = ia; inner <= ib; ++inner) {
                    const size_t q = horizontal_first ? (size_t)(h + outer) * W + (w + inner)
                                                      : (size_t)(h + inner) * W + (w + outer);
                    if (disp[q] >= 0) { vote++; hist[disp[q]] += 1; }
                }
            }
            if (vote <= ORC_VOTING_THRESH) {
                tmp[p] = disp[p];
            } else {
                int best = disp[p];
                float ratio_max = 0;
                for (int d = 0; d < Dn; ++d) {
                    float ratio = hist[d] / (float)vote;
                    if (ratio > ratio_max) {
                        ratio_max = ratio;
                        best = (ratio_max > ORC_VOTING_RATIO) ? d : best;
                    }
                    hist[d] = 0;
                }
                tmp[p] = best;
            }
        }
    memcpy(disp, tmp, npx * sizeof(int32_t));
    free(tmp); free(hist);
}

/* ----------------------------------------------------------------------- a12 */

/* source/ADCensus.cpp:1161-1239 */
void orc_proper_interpolation(int32_t* disp, int H, int W, const uint8_t* left_img)
{
    static const int dirW[16] = {0, 2, 2, 2, 0, -2, -2, -2, 1, 2, 2, 1, -1, -2, -2, -1};
    static const int dirH[16] = {2, 2, 0, -2, -2, -2, 0, 2, 2, 1, -1, -2, -2, -1, 1, 2};
    const size_t npx = (size_t)H * W;
    int32_t* tmp = (int32_t*)malloc(npx * sizeof(int32_t));
#pragma omp parallel for schedule(static)
    for (int h = 0; h < H; ++h)
        for (int w = 0; w < W; ++w) {
            const size_t p = (size_t)h * W + w;
            if (disp[p] >= 0) { tmp[p] = disp[p]; continue; }
            int nd[16], nf[16];
            for (int k = 0; k < 16; ++k) { nd[k] = disp[p]; nf[k] = -1; }
            for (int k = 0; k < 16; ++k) {
                int hD = h, wD = w, inside = 1, got = 0;
                for (int s = 0; s < ORC_MAX_SEARCH_DEPTH && inside && !got; ++s) {
                    if (s % 2 == 0) { hD += dirH[k] / 2; wD += dirW[k] / 2; }
                    else { hD += dirH[k] - dirH[k] / 2; wD += dirW[k] - dirW[k] / 2; }
                    inside = hD >= 0 && hD < H && wD >= 0 && wD < W;
                    if (inside && disp[(size_t)hD * W + wD] >= 0) {
                        nd[k] = disp[(size_t)hD * W + wD];
                        nf[k] = color_diff(left_img + p * 3, left_img + ((size_t)hD * W + wD) * 3);
                        got = 1;
                    }
                }
            }
            if (disp[p] == ORC_OCCLUSION) {
                int md = nd[0];
                for (int k = 1; k < 16; ++k)
                    if (md > nd[k]) md = nd[k];
                tmp[p] = md;
            } else {
                int md = nd[0], mf = nf[0];
                for (int k = 1; k < 16; ++k)
                    if (mf < 0 || (mf > nf[k] && nf[k] > 0)) { md = nd[k]; mf = nf[k]; }
                tmp[p] = md;
            }
        }
    memcpy(disp, tmp, npx * sizeof(int32_t));
    free(tmp);
}

/* ----------------------------------------------------------------------- a13 */

/* source/ADCensus.cpp:1241-1342.  convertDisp2Gray wraps with (uchar) (:1249). */
void orc_discontinuity_adjustment(int32_t* disp, int H, int W, int Dn, const float* vol, uint8_t* edges_out)
{
    static const int dH[8] = {-1, 1, -1, 1, -1, 1, 0, 0};
    static const int dW[8] = {-1, 1, 0, 0, 1, -1, -1, 1};
    const size_t npx = (size_t)H * W;
    uint8_t* gray = (uint8_t*)malloc(npx);
    uint8_t* blurred = (uint8_t*)malloc(npx);
    uint8_t* E = (uint8_t*)malloc(npx);
    int32_t* tmp = (int32_t*)malloc(npx * sizeof(int32_t));
    for (size_t i = 0; i < npx; ++i) gray[i] = disp[i] < 0 ? 0 : (uint8_t)disp[i];
    cvp_equalize_hist(gray, gray, H, W);
    cvp_blur3x3(gray, blurred, H, W);
    cvp_canny3(blurred, E, H, W, ORC_CANNY_LOW, ORC_CANNY_HIGH);
    if (edges_out) memcpy(edges_out, E, npx);
    memcpy(tmp, disp, npx * sizeof(int32_t));
#define EDG(yy, xx) (E[(size_t)(yy) * W + (xx)] != 0)
    for (int h = 1; h < H - 1; ++h)
        for (int w = 1; w < W - 1; ++w) {
            if (!EDG(h, w)) continue;
            int dir = -1;
            if (EDG(h - 1, w - 1) && EDG(h + 1, w + 1)) dir = 0;
            else if (EDG(h - 1, w + 1) && EDG(h + 1, w - 1)) dir = 4;
            else if (EDG(h - 1, w) || EDG(h + 1, w)) {
                if (EDG(h - 1, w - 1) || EDG(h - 1, w) || EDG(h - 1, w + 1))
                    if (EDG(h + 1, w - 1) || EDG(h + 1, w) || EDG(h + 1, w + 1)) dir = 2;
            } else {
                if (EDG(h - 1, w - 1) || EDG(h, w - 1) || EDG(h + 1, w - 1))
                    if (EDG(h - 1, w + 1) || EDG(h, w + 1) || EDG(h + 1, w + 1)) dir = 6;
            }
            if (dir == -1) continue;
            const size_t p = (size_t)h * W + w;
            int d = disp[p];
            dir = (dir + 4) % 8;
            if (d >= 0) {
                float cost = vol[p * Dn + d];
                const size_t p1 = (size_t)(h + dH[dir]) * W + (w + dW[dir]);
                const size_t p2 = (size_t)(h + dH[dir + 1]) * W + (w + dW[dir + 1]);
                const int d1 = disp[p1], d2 = disp[p2];
                const float c1 = d1 >= 0 ? vol[p1 * Dn + d1] : -1;
                const float c2 = d2 >= 0 ? vol[p2 * Dn + d2] : -1;
                if (c1 != -1 && c1 < cost) { d = d1; cost = c1; }
                if (c2 != -1 && c2 < cost) { d = d2; }
            }
            tmp[p] = d;
        }
#undef EDG
    memcpy(disp, tmp, npx * sizeof(int32_t));
    free(gray); free(blurred); free(E); free(tmp);
}

/* ----------------------------------------------------------------------- a14 */

/* source/ADCensus.cpp:1344-1374, minD = 0, maxD = Dn-1 */
void orc_subpixel(const int32_t* disp, int H, int W, int Dn, const float* vol, float* out)
{
    const size_t npx = (size_t)H * W;
    const int maxD = Dn - 1;
    float* tmp = (float*)malloc(npx * sizeof(float));
    for (size_t p = 0; p < npx; ++p) {
        const int d = disp[p];
        float f = (float)d;
        if (d > 0 && d < maxD) {
            const float cost = vol[p * Dn + d];
            const float cp = vol[p * Dn + d + 1];
            const float cm = vol[p * Dn + d - 1];
            const float diff = (cp - cm) / (2 * (cp + cm - 2 * cost));
            if (diff > -1 && diff < 1) f -= diff;
        }
        tmp[p] = f;
    }
    cvp_median3x3_f32(tmp, out, H, W);
    free(tmp);
}

/* ------------------------------------------------------------------ a1 / a15 */

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}
static void tapf(float* dst, const float* src, size_t n) { if (dst) memcpy(dst, src, n * sizeof(float)); }
static void tapi(int32_t* dst, const int32_t* src, size_t n) { if (dst) memcpy(dst, src, n * sizeof(int32_t)); }

/* ADCensus::compute (source/ADCensus.cpp:330-407) -> costInitialize, costAggregate,
 * scanlineOptimize, multiOptimize (:1376-1392). */
int orc_adcensus(const uint8_t* left, const uint8_t* right, int H, int W, int maxD, OrcTaps* t)
{
    OrcTaps none;
    if (!t) { memset(&none, 0, sizeof none); t = &none; }
    if (H < ORC_CENSUS_H || W < ORC_CENSUS_W || maxD < 1) return -1;
    const int Dn = maxD + 1;
    const size_t npx = (size_t)H * W, ncell = npx * Dn;
    float* vol[2] = {(float*)malloc(ncell * sizeof(float)), (float*)malloc(ncell * sizeof(float))};
    int32_t* arms[2][4];
    for (int k = 0; k < 2; ++k)
        for (int a = 0; a < 4; ++a) arms[k][a] = (int32_t*)malloc(npx * sizeof(int32_t));
    int32_t* dl = (int32_t*)malloc(npx * sizeof(int32_t));
    int32_t* dr = (int32_t*)malloc(npx * sizeof(int32_t));
    int32_t* disp = (int32_t*)malloc(npx * sizeof(int32_t));
    float* fin = (float*)malloc(npx * sizeof(float));
    if (!vol[0] || !vol[1] || !fin) return -2;

    double t0 = now_s();
    orc_cost_init(left, right, H, W, Dn, vol[0], vol[1]);
    double t1 = now_s();
    tapf(t->vol_init[0], vol[0], ncell); tapf(t->vol_init[1], vol[1], ncell);

    double t2 = now_s();
    orc_arms(left, H, W, arms[0][0], arms[0][1], arms[0][2], arms[0][3]);
    orc_arms(right, H, W, arms[1][0], arms[1][1], arms[1][2], arms[1][3]);
    for (int k = 0; k < 2; ++k) orc_aggregate(vol[k], H, W, Dn, arms[k][0], arms[k][1], arms[k][2], arms[k][3]);
    double t3 = now_s();
    tapf(t->vol_agg[0], vol[0], ncell); tapf(t->vol_agg[1], vol[1], ncell);
    for (int k = 0; k < 2; ++k)
        for (int a = 0; a < 4; ++a) tapi(t->arms[k][a], arms[k][a], npx);

    double t4 = now_s();
    orc_scanline(vol[0], H, W, Dn, left, right, 0);
    orc_scanline(vol[1], H, W, Dn, right, left, 1);
    double t5 = now_s();
    tapf(t->vol_scan[0], vol[0], ncell); tapf(t->vol_scan[1], vol[1], ncell);

    double t6 = now_s();
    orc_wta(vol[0], H, W, Dn, dl);
    orc_wta(vol[1], H, W, Dn, dr);
    tapi(t->wta[0], dl, npx); tapi(t->wta[1], dr, npx);
    orc_lrc(dl, dr, H, W, maxD, disp);
    tapi(t->lrc, disp, npx);
    int hf = 0;
    for (int i = 0; i < 5; ++i) {
        orc_region_voting(disp, H, W, Dn, arms[0][0], arms[0][1], arms[0][2], arms[0][3], hf);
        tapi(t->vote[i], disp, npx);
        hf = !hf;
    }
    orc_proper_interpolation(disp, H, W, left);
    tapi(t->interp, disp, npx);
    orc_discontinuity_adjustment(disp, H, W, Dn, vol[0], NULL);
    tapi(t->discont, disp, npx);
    orc_subpixel(disp, H, W, Dn, vol[0], fin);
    double t7 = now_s();
    tapf(t->final_disp, fin, npx);
    t->t_init = t1 - t0; t->t_agg = t3 - t2; t->t_scan = t5 - t4; t->t_multi = t7 - t6;

    free(vol[0]); free(vol[1]);
    for (int k = 0; k < 2; ++k)
        for (int a = 0; a < 4; ++a) free(arms[k][a]);
    free(dl); free(dr); free(disp); free(fin);
    return 0;
}

int orc_omp_max_threads(void) { return omp_get_max_threads(); }
void orc_omp_set_num_threads(int n) { if (n > 0) omp_set_num_threads(n); }
