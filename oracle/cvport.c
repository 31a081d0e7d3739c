/*
 * oracle/cvport.c -- TEST INFRASTRUCTURE (oracle), not product code.
 * See cvport.h for scope, call sites and how these restatements are pinned.
 */
#include "cvport.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline int reflect101(int p, int n)
{
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = p < 0 ? -p : 2 * n - 2 - p;
    return p;
}

/* OpenCV imgproc/histogram.cpp equalizeHist: LUT from the cumulative histogram,
 * scale = 255.f / (total - hist[first]); lut = saturate_cast<uchar>(sum * scale). */
void cvp_equalize_hist(const uint8_t* src, uint8_t* dst, int H, int W)
{
    int hist[256] = {0};
    int lut[256] = {0};
    const long total = (long)H * W;
    for (long k = 0; k < total; ++k) hist[src[k]]++;
    int i = 0;
    while (!hist[i]) ++i;
    if (hist[i] == total) {
        memset(dst, i, (size_t)total);
        return;
    }
    float scale = (256 - 1.f) / (float)(total - hist[i]);
    int sum = 0;
    for (lut[i++] = 0; i < 256; ++i) {
        sum += hist[i];
        long r = lrintf((float)sum * scale);
        lut[i] = (int)(r < 0 ? 0 : (r > 255 ? 255 : r));
    }
    for (long k = 0; k < total; ++k) dst[k] = (uint8_t)lut[src[k]];
}

/* OpenCV boxFilter, normalize=true, 8U: integer window sum then
 * saturate_cast<uchar>(sum * (1.0/9)) (round half to even; no exact .5 exists). */
void cvp_blur3x3(const uint8_t* src, uint8_t* dst, int H, int W)
{
    const double scale = 1.0 / 9.0;
    for (int y = 0; y < H; ++y) {
        for (int x = 0; x < W; ++x) {
            int s = 0;
            for (int dy = -1; dy <= 1; ++dy) {
                const uint8_t* row = src + (size_t)reflect101(y + dy, H) * W;
                for (int dx = -1; dx <= 1; ++dx) s += row[reflect101(x + dx, W)];
            }
            long r = lrint((double)s * scale);
            dst[(size_t)y * W + x] = (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
        }
    }
}

/* OpenCV imgproc/canny.cpp, aperture 3, L1 magnitude: Sobel (BORDER_REPLICATE),
 * zero-padded magnitude, fixed-point sector test (TG22 = 13573, shift 15),
 * double threshold, 8-connected hysteresis. */
void cvp_canny3(const uint8_t* src, uint8_t* dst, int H, int W, int low, int high)
{
    if (low > high) { int t = low; low = high; high = t; }
    const int MW = W + 2;
    int* mag = (int*)calloc((size_t)(H + 2) * MW, sizeof(int));
    short* gx = (short*)malloc((size_t)H * W * sizeof(short));
    short* gy = (short*)malloc((size_t)H * W * sizeof(short));
    uint8_t* map = (uint8_t*)malloc((size_t)H * W); /* 0 weak candidate, 1 none, 2 edge */
    long* stack = (long*)malloc((size_t)H * W * sizeof(long));
    long sp = 0;
#define SRC(yy, xx) ((int)src[(size_t)clampi((yy), 0, H - 1) * W + clampi((xx), 0, W - 1)])
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            int dx = (SRC(y - 1, x + 1) - SRC(y - 1, x - 1)) + 2 * (SRC(y, x + 1) - SRC(y, x - 1)) +
                     (SRC(y + 1, x + 1) - SRC(y + 1, x - 1));
            int dy = (SRC(y + 1, x - 1) - SRC(y - 1, x - 1)) + 2 * (SRC(y + 1, x) - SRC(y - 1, x)) +
                     (SRC(y + 1, x + 1) - SRC(y - 1, x + 1));
            gx[(size_t)y * W + x] = (short)dx;
            gy[(size_t)y * W + x] = (short)dy;
            mag[(size_t)(y + 1) * MW + x + 1] = abs(dx) + abs(dy);
        }
#undef SRC
    const int TG22 = 13573;
    for (int y = 0; y < H; ++y) {
        const int* mp = mag + (size_t)y * MW + 1;       /* row y-1 */
        const int* mc = mag + (size_t)(y + 1) * MW + 1; /* row y   */
        const int* mn = mag + (size_t)(y + 2) * MW + 1; /* row y+1 */
        for (int x = 0; x < W; ++x) {
            int m = mc[x];
            uint8_t v = 1;
            if (m > low) {
                int xs = gx[(size_t)y * W + x], ys = gy[(size_t)y * W + x];
                int ax = abs(xs), ay = abs(ys) << 15;
                int tg22x = ax * TG22;
                int is_max;
                if (ay < tg22x) {
                    is_max = m > mc[x - 1] && m >= mc[x + 1];
                } else {
                    int tg67x = tg22x + (ax << 16);
                    if (ay > tg67x) {
                        is_max = m > mp[x] && m >= mn[x];
                    } else {
                        int s = (xs ^ ys) < 0 ? -1 : 1;
                        is_max = m > mp[x - s] && m > mn[x + s];
                    }
                }
                if (is_max) {
                    if (m > high) { v = 2; stack[sp++] = (long)y * W + x; }
                    else v = 0;
                }
            }
            map[(size_t)y * W + x] = v;
        }
    }
    while (sp > 0) {
        long p = stack[--sp];
        int y = (int)(p / W), x = (int)(p % W);
        for (int dy = -1; dy <= 1; ++dy)
            for (int dx = -1; dx <= 1; ++dx) {
                int yy = y + dy, xx = x + dx;
                if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
                size_t q = (size_t)yy * W + xx;
                if (map[q] == 0) { map[q] = 2; stack[sp++] = (long)q; }
            }
    }
    for (size_t k = 0; k < (size_t)H * W; ++k) dst[k] = map[k] == 2 ? 255 : 0;
    free(mag); free(gx); free(gy); free(map); free(stack);
}

static inline void sort2f(float* a, float* b)
{
    float lo = *a < *b ? *a : *b, hi = *a < *b ? *b : *a;
    *a = lo; *b = hi;
}

/* OpenCV medianBlur ksize 3, CV_32F: exact median of the 3x3 window, BORDER_REPLICATE. */
void cvp_median3x3_f32(const float* src, float* dst, int H, int W)
{
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            float p[9];
            int k = 0;
            for (int dy = -1; dy <= 1; ++dy)
                for (int dx = -1; dx <= 1; ++dx)
                    p[k++] = src[(size_t)clampi(y + dy, 0, H - 1) * W + clampi(x + dx, 0, W - 1)];
            /* 19-exchange median-of-9 network */
            sort2f(&p[1], &p[2]); sort2f(&p[4], &p[5]); sort2f(&p[7], &p[8]);
            sort2f(&p[0], &p[1]); sort2f(&p[3], &p[4]); sort2f(&p[6], &p[7]);
            sort2f(&p[1], &p[2]); sort2f(&p[4], &p[5]); sort2f(&p[7], &p[8]);
            sort2f(&p[0], &p[3]); sort2f(&p[5], &p[8]); sort2f(&p[4], &p[7]);
            sort2f(&p[3], &p[6]); sort2f(&p[1], &p[4]); sort2f(&p[2], &p[5]);
            sort2f(&p[4], &p[7]); sort2f(&p[4], &p[2]); sort2f(&p[6], &p[4]);
            sort2f(&p[4], &p[2]);
            dst[(size_t)y * W + x] = p[4];
        }
}

/* OpenCV imgproc/imgwarp.cpp remapBilinear<FixedPtCast<int,uchar,15>> with the
 * 32x32 bilinear table: weights a*b*32 (a in {32-fx,fx}, b in {32-fy,fy}) sum to
 * 32768, so dst = (sum(a*b*src) + 512) >> 10.  Out-of-image taps read 0. */
void cvp_remap_bilinear_8uc3_fixed(const uint8_t* src, int sH, int sW, int sstep,
                                   const int16_t* map1, const uint16_t* map2,
                                   uint8_t* dst, int H, int W)
{
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            size_t k = (size_t)y * W + x;
            int sx = map1[2 * k], sy = map1[2 * k + 1];
            int fx = map2[k] & 31, fy = (map2[k] >> 5) & 31;
            int wx[2] = {32 - fx, fx}, wy[2] = {32 - fy, fy};
            for (int c = 0; c < 3; ++c) {
                int acc = 0;
                for (int j = 0; j < 2; ++j)
                    for (int i = 0; i < 2; ++i) {
                        int xx = sx + i, yy = sy + j;
                        int v = (xx >= 0 && xx < sW && yy >= 0 && yy < sH)
                                    ? src[(size_t)yy * sstep + 3 * xx + c] : 0;
                        acc += wx[i] * wy[j] * v;
                    }
                dst[3 * k + c] = (uint8_t)((acc + 512) >> 10);
            }
        }
}

static inline int16_t sat_s16(int v) { return (int16_t)(v < -32768 ? -32768 : (v > 32767 ? 32767 : v)); }

/* OpenCV convertMaps (CV_32FC1 x2 -> CV_16SC2 + CV_16UC1): ix = cvRound(x*32). */
void cvp_convert_maps_f32(const float* mx, const float* my, int H, int W,
                          int16_t* map1, uint16_t* map2)
{
    for (size_t k = 0; k < (size_t)H * W; ++k) {
        int ix = (int)lrintf(mx[k] * 32.f);
        int iy = (int)lrintf(my[k] * 32.f);
        map1[2 * k] = sat_s16(ix >> 5);
        map1[2 * k + 1] = sat_s16(iy >> 5);
        map2[k] = (uint16_t)((iy & 31) * 32 + (ix & 31));
    }
}
