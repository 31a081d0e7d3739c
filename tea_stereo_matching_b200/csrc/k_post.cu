// k_post.cu -- disparity refinement chain (multiOptimize, reference
// source/ADCensus.cpp:1376-1392): WTA, left-right check, region voting x5 (including
// the reference's histogram leak), proper interpolation, depth-discontinuity
// adjustment (equalizeHist -> blur -> Canny -> edge fix-up), sub-pixel + 3x3 median.
// All integer stages are bit-exact restatements; see each kernel for the cited lines.
#include "tsm_common.cuh"
#include <cooperative_groups.h>
#include <float.h>
#include <limits.h>

namespace cg = cooperative_groups;

namespace tsm {

// =========================================================================== a9
// cost2disparity (ADCensus.cpp:1394-1413): first strict minimum over the PLANES minD .. maxD - minD = minD .. Dn - 1; the plane
// index is what the reference reports as the disparity.
__global__ void __launch_bounds__(256) k_wta(Vol vol, Dims dm, int32_t* __restrict__ disp)
{
    const size_t npx = dm.npx();
    const int Dn = dm.Dn;
    const size_t p = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (p >= npx) return;
    float best = FLT_MAX;
    int bd = INT_MAX;
    for (int d = lane; d < Dn; d += 32) {
        if (d < dm.minD) continue;
        const float v = *cell_ptr(vol, dm, p, d);
        if (best > v) { best = v; bd = d; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int od = __shfl_xor_sync(0xffffffffu, bd, o);
        if (ov < best || (ov == best && od < bd)) { best = ov; bd = od; }
    }
    if (lane == 0) disp[p] = bd == INT_MAX ? 0 : bd;
}

void wta(const Launcher& L, const Dims& d, const Vol& vol, int32_t* disp)
{
    const size_t npx = d.npx();
    k_wta<<<(unsigned)((npx + 7) / 8), 256, 0, L.stream>>>(vol, d, disp);
    L.count(1);
}

// ---- dense <-> split volume copies for the parity taps ----
__global__ void k_volume_copy(Vol vol, Dims dm, float* dense, int to_dense)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t n = dm.npx() * dm.Dn;
    if (i >= n) return;
    const size_t p = i / dm.Dn;
    const int d = (int)(i % dm.Dn);
    float* c = cell_ptr(vol, dm, p, d);
    if (to_dense) dense[i] = *c;
    else *c = dense[i];
}
void volume_gather(const Launcher& L, const Dims& d, const Vol& vol, float* dense)
{
    const size_t n = d.npx() * d.Dn;
    k_volume_copy<<<(unsigned)((n + 255) / 256), 256, 0, L.stream>>>(vol, d, dense, 1);
    L.count(1);
}
void volume_scatter(const Launcher& L, const Dims& d, const float* dense, const Vol& vol)
{
    const size_t n = d.npx() * d.Dn;
    k_volume_copy<<<(unsigned)((n + 255) / 256), 256, 0, L.stream>>>(vol, d, const_cast<float*>(dense), 0);
    L.count(1);
}

// ========================================================================== a10
// outlierElimination (ADCensus.cpp:1013-1044), dispTolerance = 0; the markers stay -1 / -2 whatever minD is (:415-416).
__global__ void k_lrc(const int32_t* __restrict__ dl, const int32_t* __restrict__ dr, int32_t* __restrict__ out, int H, int W, int minD,
                      int maxD)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const int32_t* rrow = dr + (size_t)y * W;
    int disp = dl[(size_t)y * W + x];
    if (x - disp < 0 || rrow[x - disp] != disp) {
        bool occlusion = true;
        const int dmax = min(maxD, x);
        for (int d = minD; d <= dmax; ++d)
            if (rrow[x - d] == d) { occlusion = false; break; }
        disp = occlusion ? kOcclusion : kMismatch;
    }
    out[(size_t)y * W + x] = disp;
}

void lrc(const Launcher& L, const Dims& d, const int32_t* dl, const int32_t* dr, int32_t* out)
{
    dim3 g((d.W + 127) / 128, d.H);
    k_lrc<<<g, 128, 0, L.stream>>>(dl, dr, out, d.H, d.W, d.minD, d.minD + d.Dn - 1);
    L.count(1);
}

// ========================================================================== a11
// regionVoting (ADCensus.cpp:1046-1159).  The reference walks pixels in raster order
// with ONE histogram that is cleared only after a high-vote outlier (:1150), so the
// votes of every low-vote outlier (vote <= 20) leak into the next high-vote outlier.
// Exact parallel form:
//   1. k_vote_prefix : prefix counts of the valid mask -> vote counts without a traversal (see below)
//   2. k_vote_pass_a : per outlier with votes ONE traversal of its region: histogram; high-vote outliers
//                      are decided from their own histogram, low-vote ones park their (<= 20) votes
//   3. k_vscan_*     : one exclusive scan over raster order -> CSR offsets of the parked votes, start of the run of
//                      low-vote outliers preceding each high-vote outlier, parked votes -> CSR payload
//   4. k_vote_pass_b : only the high-vote outliers whose slice [start, off) is not empty are redone with
//                      own region + leaked slice (first arg-max, ratio test (float)h/(float)vote > 0.4f)
constexpr unsigned kVoteHigh = 0x80000000u;  // scan element of a high-vote outlier
// Cross region of p (arms of the LEFT view): horizontal_first: rows y-up..y+down, each
// with its own left/right arm; else columns x-left..x+right, each with its own up/down arm.

// ---- vote counts without a traversal, rows worth walking -----------------------------------------------------------------------
// A third to two thirds of the outliers sit inside blobs of outliers (the left border band, occlusions): their cross region holds
// no valid pixel, in every one of the five iterations (measured at C3: 60 / 50 / 46 / 35 / 30 % of the outliers, 34 - 57 % on the
// 0600 pair).  The vote COUNT is separable -- the region is a union of segments, one per row (horizontal first) or one per column
// (vertical first) -- so it is a sum of <= 67 prefix-count differences: one strip-local prefix count of the valid mask per
// iteration along the segment direction (k_vote_prefix), two or three byte loads per segment.  Outliers without a vote are never
// traversed.  Horizontal-first regions additionally walk only the rows whose segment holds a valid pixel (the walk is row by row
// there); for vertical-first regions the rows of the bounding box were tried as a row filter and lost on the real pairs (C2
// 3.56 -> 3.83 ms: next to valid areas nearly every row of the box holds some valid pixel), so they are walked in full.
// Strips of 128 pixels (a segment is at most 67 long: it crosses at most one strip boundary) keep the counts in one byte.
constexpr int VP_STRIP = 128;
template <bool VERT>
__global__ void __launch_bounds__(256) k_vote_prefix(const int32_t* __restrict__ disp, uint8_t* __restrict__ pre, int H, int W, int minD)
{
    if (VERT) {
        // counts along y: a thread owns (column, strip of 128 rows)
        const int x = blockIdx.x * blockDim.x + threadIdx.x;
        if (x >= W) return;
        const int y0 = blockIdx.y * VP_STRIP, y1 = min(y0 + VP_STRIP, H);
        int run = 0;
#pragma unroll 8
        for (int y = y0; y < y1; ++y) {
            run += disp[(size_t)y * W + x] >= minD ? 1 : 0;
            pre[(size_t)y * W + x] = (uint8_t)run;
        }
        return;
    }
    // counts along x: a warp owns (row, strip of 128 columns), a lane 4 consecutive pixels
    const int lane = threadIdx.x & 31;
    const int nstrip = (W + VP_STRIP - 1) / VP_STRIP;
    const int item = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (item >= H * nstrip) return;
    const int y = item / nstrip, x0 = (item % nstrip) * VP_STRIP + 4 * lane;
    int v[4], sum = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        v[k] = (x0 + k < W && disp[(size_t)y * W + x0 + k] >= minD) ? 1 : 0;
        sum += v[k];
    }
    int inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    int run = inc - sum;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        run += v[k];
        if (x0 + k < W) pre[(size_t)y * W + x0 + k] = (uint8_t)run;
    }
}

// valid pixels at positions q0..q1 (inclusive, q0 <= q1 <= q0 + 127) of the line whose strip-local prefix counts start at `line`
// with element stride `stride`
__device__ __forceinline__ int vp_count(const uint8_t* __restrict__ line, size_t stride, int q0, int q1)
{
    int r = line[(size_t)q1 * stride];
    if (q0 & (VP_STRIP - 1)) r -= line[(size_t)(q0 - 1) * stride];
    if ((q0 ^ q1) & ~(VP_STRIP - 1)) r += line[(size_t)(q0 | (VP_STRIP - 1)) * stride];  // the strip q0 lies in, up to its end
    return r;
}

struct RegionRows {
    unsigned mask[3];  // horizontal first: bit b of mask[ch] = the segment of row y + first + 32 ch + b holds a valid pixel
    unsigned seg[3];   // horizontal first, lane b: left | right << 8 arm of that row (the walk gets them by shuffle, not by a dependent load)
    int first;         // first row of the region, relative to y
    int cnt;           // the vote count of the region
};
template <bool HF>
__device__ __forceinline__ RegionRows region_rows(const uint8_t* __restrict__ pre, const uchar4* __restrict__ arms, int W, size_t p, int lane)
{
    const uchar4 a = arms[p];
    const int y = (int)(p / W), x = (int)(p - (size_t)y * W);
    RegionRows R;
    R.first = -(int)a.x;
    int cnt = 0;
    // segments: rows y - up .. y + down with their own left / right arm (horizontal first, prefix counts along x), or
    // columns x - left .. x + right with their own up / down arm (vertical first, prefix counts along y)
    const int n = HF ? (int)a.x + (int)a.y + 1 : (int)a.z + (int)a.w + 1;  // <= 67
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        const int o = 32 * ch + lane;
        int cj = 0;
        R.seg[ch] = 0;
        if (o < n) {
            if (HF) {
                const int c = y - (int)a.x + o;
                const uchar4 ac = arms[(size_t)c * W + x];
                R.seg[ch] = (unsigned)ac.z | ((unsigned)ac.w << 8);
                cj = vp_count(pre + (size_t)c * W, 1, x - (int)ac.z, x + (int)ac.w);
            } else {
                const int c = x - (int)a.z + o;
                const uchar4 ac = arms[(size_t)y * W + c];
                cj = vp_count(pre + c, (size_t)W, y - (int)ac.x, y + (int)ac.y);
            }
        }
        R.mask[ch] = __ballot_sync(0xffffffffu, cj > 0);
        cnt += cj;
    }
    R.cnt = __reduce_add_sync(0xffffffffu, cnt);
    return R;
}

// Calls f(valid, value) warp-synchronously for every pixel of the cross region of p (horizontal first: of its rows in R).
// Lanes always run along x so the disparity reads are coalesced.
template <bool HF, typename F>
__device__ __forceinline__ void for_each_region(const int32_t* __restrict__ disp, const uchar4* __restrict__ arms, int W,
                                                size_t p, int lane, int minD, const RegionRows& R, F f)
{
    const uchar4 a = arms[p];
    if (HF) {
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            for (unsigned m = R.mask[ch]; m; m &= m - 1) {
                const int b = __ffs(m) - 1;
                const unsigned sg = __shfl_sync(0xffffffffu, R.seg[ch], b);
                const int left = (int)(sg & 0xffu), right = (int)(sg >> 8);
                const int32_t* q = disp + p + (ptrdiff_t)(R.first + 32 * ch + b) * W;
                for (int i0 = -left; i0 <= right; i0 += 32) {
                    const int i = i0 + lane;
                    const bool in = i <= right;
                    const int v = in ? q[i] : -1;
                    f(in && v >= minD, v - minD);  // valid: disp >= minD; histogram bin disp - minD (:1127-1131)
                }
            }
        }
    } else {
        const int no = (int)a.z + (int)a.w + 1;
        if (no <= 16) {
            // A narrow region (the mean arm is 9 pixels) leaves most lanes of a row idle: lane groups of w = 2, 4, 8 or 16 lanes
            // take 32 / w rows per iteration instead of one.
            const int sh = no <= 2 ? 1 : no <= 4 ? 2 : no <= 8 ? 3 : 4;
            const int col = lane & ((1 << sh) - 1), h = lane >> sh, rpi = 32 >> sh;
            const bool oin = col < no;
            const size_t c = p + (oin ? col - (int)a.z : 0);
            const uchar4 ac = arms[c];
            const int up = oin ? (int)ac.x : -1, down = oin ? (int)ac.y : -1;  // -1: no row matches
            const int mup = __reduce_max_sync(0xffffffffu, up), mdown = __reduce_max_sync(0xffffffffu, down);
            const int32_t* q = disp + c + (ptrdiff_t)(h - mup) * W;  // running pointer: the 64-bit index arithmetic is paid once
            const ptrdiff_t qstep = (ptrdiff_t)rpi * W;
            for (int i = h - mup; i <= mdown + h; i += rpi, q += qstep) {
                const bool in = i >= -up && i <= down;
                const int v = in ? *q : -1;
                f(in && v >= minD, v - minD);  // valid: disp >= minD; histogram bin disp - minD (:1127-1131)
            }
            return;
        }
        for (int o0 = 0; o0 < no; o0 += 32) {
            const int o = o0 + lane - (int)a.z;
            const bool oin = o <= (int)a.w;
            const size_t c = p + (oin ? o : 0);
            const uchar4 ac = arms[c];
            int up = oin ? (int)ac.x : 0, down = oin ? (int)ac.y : 0;
            int mup = up, mdown = down;
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) {
                mup = max(mup, __shfl_xor_sync(0xffffffffu, mup, s));
                mdown = max(mdown, __shfl_xor_sync(0xffffffffu, mdown, s));
            }
            const int32_t* q = disp + c - (ptrdiff_t)mup * W;
            for (int i = -mup; i <= mdown; ++i, q += W) {
                const bool in = oin && i >= -up && i <= down;
                const int v = in ? *q : -1;
                f(in && v >= minD, v - minD);  // valid: disp >= minD; histogram bin disp - minD (:1127-1131)
            }
        }
    }
}

// A CTA owns a tile of 256 consecutive pixels: every thread classifies its own pixel, the
// pixels that need region work are compacted into a shared list and the 8 warps then take
// them one at a time (most pixels need none, so a warp per pixel would mostly launch to exit).
constexpr int VOTE_TILE = 256, VOTE_WARPS = 8;
__device__ __forceinline__ int tile_compact(bool want, int local, int* list, int* count)
{
    // order inside the list is irrelevant; returns the number of entries after the barrier
    const unsigned b = __ballot_sync(0xffffffffu, want);
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == 0 && b) base = atomicAdd(count, __popc(b));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (want) list[base + __popc(b & ((1u << lane) - 1u))] = local;
    __syncthreads();
    return *count;
}

// Adds the warp's votes to its shared-memory histogram: equal values are grouped first (match.any), so a
// region full of one disparity costs one shared atomic instead of a 32-way serialised one.
__device__ __forceinline__ void hist_add(uint32_t hist_s, bool valid, int v, int lane)
{
    // hist_s: the histogram's 32-bit shared-memory address (a generic pointer costs six instructions of address arithmetic per call)
    const unsigned act = __ballot_sync(0xffffffffu, valid);
    if (!act) return;
    if (valid) {
        const unsigned peers = __match_any_sync(act, v);
        if (lane == __ffs(peers) - 1)
            asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(hist_s + 4u * (uint32_t)v), "r"((uint32_t)__popc(peers)) : "memory");
    }
}

// First arg-max over d ascending + the ratio test of ADCensus.cpp:1138-1152.
__device__ __forceinline__ int vote_decide(const int* hist, int Dn, int nv, int dp, int lane, int minD)
{
    int best = 0, bd = INT_MAX;
    for (int d = lane; d < Dn; d += 32) {
        const int h = hist[d];
        if (h > best) { best = h; bd = d; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const int oh = __shfl_xor_sync(0xffffffffu, best, o);
        const int od = __shfl_xor_sync(0xffffffffu, bd, o);
        if (oh > best || (oh == best && od < bd)) { best = oh; bd = od; }
    }
    const float ratio = __fdiv_rn((float)best, (float)nv);  // hist[d] / (float)vote, :1144
    return (best > 0 && ratio > kVotingRatio) ? bd + minD : dp;
}

// Pass A: ONE traversal of the cross region per outlier builds its histogram and its vote count.
//   vote > 20 : decided right here from the own histogram -- final unless low-vote outliers precede it in raster
//               order since the last high-vote one (their votes leak in, pass B redoes exactly those pixels);
//   0 < vote <= 20 : pixel unchanged; its votes are parked in stash[p][0..vote) for the leak;
//   vote == 0, valid pixels : unchanged.
template <bool HF>
__global__ void __launch_bounds__(VOTE_TILE)
k_vote_pass_a(const int32_t* __restrict__ disp, const uchar4* __restrict__ arms, const uint8_t* __restrict__ pre,
              int32_t* __restrict__ vote, int32_t* __restrict__ lowcnt, uint16_t* __restrict__ stash, int32_t* __restrict__ out,
              size_t npx, int W, int Dn, int minD)
{
    extern __shared__ int hist_all[];  // [VOTE_WARPS][Dn]
    __shared__ int list[VOTE_TILE];
    __shared__ int count;
    if (threadIdx.x == 0) count = 0;
    __syncthreads();
    const size_t p0 = (size_t)blockIdx.x * VOTE_TILE, pt = p0 + threadIdx.x;
    bool outlier = false;
    if (pt < npx) {
        const int dp = disp[pt];
        outlier = dp < minD;
        if (!outlier) { vote[pt] = 0; lowcnt[pt] = 0; out[pt] = dp; }
    }
    const int n = tile_compact(outlier, threadIdx.x, list, &count);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int* hist = hist_all + warp * Dn;
    const uint32_t hist_s = (uint32_t)__cvta_generic_to_shared(hist);
    for (int i = warp; i < n; i += VOTE_WARPS) {
        const size_t p = p0 + list[i];
        const int dp = disp[p];
        const RegionRows R = region_rows<HF>(pre, arms, W, p, lane);
        if (R.cnt == 0) {  // empty region: nothing to vote with, nothing to park
            if (lane == 0) {
                vote[p] = 0;
                lowcnt[p] = 0;
                out[p] = dp;
            }
            continue;
        }
        for (int d = lane; d < Dn; d += 32) hist[d] = 0;
        __syncwarp();
        const int cnt = R.cnt;
        for_each_region<HF>(disp, arms, W, p, lane, minD, R, [&](bool valid, int v) { hist_add(hist_s, valid, v, lane); });
        __syncwarp();
        int res = dp;
        if (cnt > kVotingThresh) {
            res = vote_decide(hist, Dn, cnt, dp, lane, minD);
        } else if (cnt > 0) {
            // park the votes (order is irrelevant for a histogram): bin d contributes hist[d] copies of d
            uint16_t* dst = stash + p * kVotingThresh;
            int base = 0;
            for (int d0 = 0; d0 < Dn; d0 += 32) {
                const int d = d0 + lane;
                const int h = d < Dn ? hist[d] : 0;
                int incl = h;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += t;
                }
                for (int k = 0; k < h; ++k) dst[base + incl - h + k] = (uint16_t)d;
                base += __shfl_sync(0xffffffffu, incl, 31);
            }
        }
        if (lane == 0) {
            vote[p] = cnt;
            lowcnt[p] = cnt <= kVotingThresh ? cnt : (int)kVoteHigh;  // scan element: parked votes, or the high-vote flag
            out[p] = res;
        }
        __syncwarp();
    }
}

// Pass B: high-vote outliers that inherit leaked votes (start < off): own region again + the slice.
template <bool HF>
__global__ void __launch_bounds__(VOTE_TILE)
k_vote_pass_b(const int32_t* __restrict__ disp, const uchar4* __restrict__ arms, const uint8_t* __restrict__ pre,
              const int32_t* __restrict__ vote, const int32_t* __restrict__ off, const int32_t* __restrict__ start,
              const uint16_t* __restrict__ flat, int32_t* __restrict__ out, size_t npx, int W, int Dn, int minD)
{
    extern __shared__ int hist_all[];  // [VOTE_WARPS][Dn]
    __shared__ int list[VOTE_TILE];
    __shared__ int count;
    if (threadIdx.x == 0) count = 0;
    __syncthreads();
    const size_t p0 = (size_t)blockIdx.x * VOTE_TILE, pt = p0 + threadIdx.x;
    bool redo = false;
    if (pt < npx) redo = disp[pt] < minD && vote[pt] > kVotingThresh && start[pt] < off[pt];
    const int n = tile_compact(redo, threadIdx.x, list, &count);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int* hist = hist_all + warp * Dn;
    const uint32_t hist_s = (uint32_t)__cvta_generic_to_shared(hist);
    for (int i = warp; i < n; i += VOTE_WARPS) {
        const size_t p = p0 + list[i];
        for (int d = lane; d < Dn; d += 32) hist[d] = 0;
        __syncwarp();
        const RegionRows R = region_rows<HF>(pre, arms, W, p, lane);
        for_each_region<HF>(disp, arms, W, p, lane, minD, R, [&](bool valid, int v) { hist_add(hist_s, valid, v, lane); });
        for (int j0 = start[p]; j0 < off[p]; j0 += 32) {  // the leak
            const int j = j0 + lane;
            const bool in = j < off[p];
            hist_add(hist_s, in, in ? (int)flat[j] : 0, lane);
        }
        __syncwarp();
        const int res = vote_decide(hist, Dn, vote[p], disp[p], lane, minD);
        if (lane == 0) out[p] = res;
        __syncwarp();
    }
}

// ---- one scan over the raster order gives both CSR arrays ------------------------------------------------------------------------
// Element of pixel p: c = its parked votes (low-vote outlier), or "high-vote outlier", or nothing.  Wanted, exclusive over q < p:
//   off[p]   = sum of c                                  -> where p's parked votes go in the CSR payload
//   start[p] = off at the LAST high-vote outlier before p -> start of the slice of votes that leak into the next high-vote outlier
// (the reference's histogram is cleared only after a high-vote outlier).  Both come from ONE scan with the associative operator
//   (sum_a, last_a) o (sum_b, last_b) = (sum_a + sum_b,  last_b >= 0 ? sum_a + last_b : last_a),        last = -1: no high-vote pixel yet,
// whose last phase also moves the parked votes into the payload: 3 launches per iteration instead of two scans, a mark and a copy
// kernel (8 launches; the small launches of the five iterations were a third of the stage's time on one stream).
constexpr int SCAN_T = 256, SCAN_I = 8, SCAN_TILE = SCAN_T * SCAN_I;
struct VS {
    int sum, last;
};
__device__ __forceinline__ VS vs_comb(VS a, VS b)
{
    VS r;
    r.sum = a.sum + b.sum;
    r.last = b.last >= 0 ? a.sum + b.last : a.last;
    return r;
}
__device__ __forceinline__ VS vs_elem(int e)
{
    VS x;
    x.sum = (unsigned)e == kVoteHigh ? 0 : e;
    x.last = (unsigned)e == kVoteHigh ? 0 : -1;
    return x;
}
__device__ __forceinline__ VS vs_shfl_up(VS v, int o)
{
    VS t;
    t.sum = __shfl_up_sync(0xffffffffu, v.sum, o);
    t.last = __shfl_up_sync(0xffffffffu, v.last, o);
    return t;
}

// exclusive scan of one VS per thread over the CTA; *total = the CTA's combined value
__device__ __forceinline__ VS block_exclusive_vscan(VS v, VS* total)
{
    __shared__ VS wsum[SCAN_T / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    VS inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const VS t = vs_shfl_up(inc, o);
        if (lane >= o) inc = vs_comb(t, inc);
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        VS w = lane < SCAN_T / 32 ? wsum[lane] : VS{0, -1};
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const VS t = vs_shfl_up(w, o);
            if (lane >= o) w = vs_comb(t, w);
        }
        if (lane < SCAN_T / 32) wsum[lane] = w;
    }
    __syncthreads();
    const VS prefix = warp > 0 ? wsum[warp - 1] : VS{0, -1};
    VS exc = vs_shfl_up(inc, 1);
    if (lane == 0) exc = VS{0, -1};
    if (total) *total = wsum[SCAN_T / 32 - 1];
    const VS r = vs_comb(prefix, exc);
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(SCAN_T) k_vscan_reduce(const int32_t* __restrict__ elem, int2* __restrict__ sums, size_t n)
{
    const size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_I;
    VS acc{0, -1};
#pragma unroll
    for (int i = 0; i < SCAN_I; ++i)
        if (base + i < n) acc = vs_comb(acc, vs_elem(elem[base + i]));
    VS total;
    block_exclusive_vscan(acc, &total);
    if (threadIdx.x == 0) sums[blockIdx.x] = make_int2(total.sum, total.last);
}

__global__ void __launch_bounds__(SCAN_T) k_vscan_sums(int2* __restrict__ sums, int nblocks)
{
    __shared__ VS carry_s;
    if (threadIdx.x == 0) carry_s = VS{0, -1};
    __syncthreads();
    for (int b0 = 0; b0 < nblocks; b0 += SCAN_T) {
        const int i = b0 + threadIdx.x;
        VS v{0, -1};
        if (i < nblocks) {
            const int2 t = sums[i];
            v = VS{t.x, t.y};
        }
        VS total;
        const VS exc = block_exclusive_vscan(v, &total);
        const VS carry = carry_s;
        if (i < nblocks) {
            const VS r = vs_comb(carry, exc);
            sums[i] = make_int2(r.sum, r.last);
        }
        __syncthreads();
        if (threadIdx.x == 0) carry_s = vs_comb(carry, total);
        __syncthreads();
    }
}

__global__ void __launch_bounds__(SCAN_T)
k_vscan_apply(const int32_t* __restrict__ elem, const int2* __restrict__ sums, int32_t* __restrict__ off, int32_t* __restrict__ start,
              const uint16_t* __restrict__ stash, uint16_t* __restrict__ flat, size_t n)
{
    const size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_I;
    int e[SCAN_I];
    VS acc{0, -1};
#pragma unroll
    for (int i = 0; i < SCAN_I; ++i) {
        e[i] = base + i < n ? elem[base + i] : 0;
        acc = vs_comb(acc, vs_elem(e[i]));
    }
    const int2 bs = sums[blockIdx.x];
    VS run = vs_comb(VS{bs.x, bs.y}, block_exclusive_vscan(acc, nullptr));
#pragma unroll
    for (int i = 0; i < SCAN_I; ++i) {
        if (base + i < n) {
            off[base + i] = run.sum;
            start[base + i] = run.last < 0 ? 0 : run.last;
            if ((unsigned)e[i] != kVoteHigh && e[i] > 0) {  // parked votes -> CSR payload, in raster order of their pixels
                const uint16_t* src = stash + (base + i) * kVotingThresh;
                uint16_t* dst = flat + run.sum;
                for (int k = 0; k < e[i]; ++k) dst[k] = src[k];
            }
        }
        run = vs_comb(run, vs_elem(e[i]));
    }
}

template <bool HF>
static void region_voting_t(const Launcher& L, const Dims& d, const int32_t* disp_in, int32_t* disp_out, const uchar4* arms,
                            const VoteScratch& s)
{
    const size_t npx = d.npx();
    const unsigned wblocks = (unsigned)((npx + VOTE_TILE - 1) / VOTE_TILE);
    const size_t smem = (size_t)VOTE_WARPS * d.Dn * sizeof(int);
    // strip-local prefix counts of the valid mask along the segment direction (both passes read them)
    if (HF) {
        const int items = d.H * ((d.W + VP_STRIP - 1) / VP_STRIP);
        k_vote_prefix<false><<<(items + 7) / 8, 256, 0, L.stream>>>(disp_in, s.pre, d.H, d.W, d.minD);
    } else {
        dim3 g((d.W + 255) / 256, (d.H + VP_STRIP - 1) / VP_STRIP);
        k_vote_prefix<true><<<g, 256, 0, L.stream>>>(disp_in, s.pre, d.H, d.W, d.minD);
    }
    k_vote_pass_a<HF><<<wblocks, VOTE_TILE, smem, L.stream>>>(disp_in, arms, s.pre, s.vote, s.lowcnt, s.stash, disp_out, npx, d.W, d.Dn, d.minD);
    L.count(2);
    {
        const int nb = (int)((npx + SCAN_TILE - 1) / SCAN_TILE);
        int2* sums = reinterpret_cast<int2*>(s.blocksums);
        k_vscan_reduce<<<nb, SCAN_T, 0, L.stream>>>(s.lowcnt, sums, npx);
        k_vscan_sums<<<1, SCAN_T, 0, L.stream>>>(sums, nb);
        k_vscan_apply<<<nb, SCAN_T, 0, L.stream>>>(s.lowcnt, sums, s.off, s.start, s.stash, s.flat, npx);
        L.count(3);
    }
    k_vote_pass_b<HF><<<wblocks, VOTE_TILE, smem, L.stream>>>(
        disp_in, arms, s.pre, s.vote, s.off, s.start, s.flat, disp_out, npx, d.W, d.Dn, d.minD);
    L.count(1);
}

void region_voting(const Launcher& L, const Dims& d, const int32_t* disp_in, int32_t* disp_out, const uchar4* arms_left,
                   bool horizontal_first, const VoteScratch& s)
{
    if (horizontal_first) region_voting_t<true>(L, d, disp_in, disp_out, arms_left, s);
    else region_voting_t<false>(L, d, disp_in, disp_out, arms_left, s);
}

// ========================================================================== a12
// properInterpolation (ADCensus.cpp:1161-1239).
__constant__ int c_dirW[16] = {0, 2, 2, 2, 0, -2, -2, -2, 1, 2, 2, 1, -1, -2, -2, -1};
__constant__ int c_dirH[16] = {2, 2, 0, -2, -2, -2, 0, 2, 2, 1, -1, -2, -2, -1, 1, 2};

__global__ void k_interpolate(const int32_t* __restrict__ disp, int32_t* __restrict__ out, const uint32_t* __restrict__ img4, int H, int W,
                              int hsi, int minD)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x, h = blockIdx.y * blockDim.y + threadIdx.y;
    if (w >= W || h >= H) return;
    const size_t p = (size_t)h * W + w;
    const int own = disp[p];
    if (own >= minD) { out[p] = own; return; }
    const uint32_t pc = img4[p];
    // occlusion: min over the 16 entries, each initialised to the pixel's own value (:1180, :1211-1216),
    // so a single direction without a hit keeps the pixel at -1
    int occ_min = INT_MAX;
    // mismatch: running pick (:1222-1231)
    int md = own, mf = -1;
#pragma unroll 1
    for (int k = 0; k < 16; ++k) {
        int hD = h, wD = w;
        bool inside = true, got = false;
        int nd = own, nf = -1;
        const int dh = c_dirH[k], dw = c_dirW[k];
        for (int s = 0; s < kMaxSearchDepth && inside && !got; ++s) {
            if ((s & 1) == 0) { hD += dh / 2; wD += dw / 2; }
            else { hD += dh - dh / 2; wD += dw - dw / 2; }
            inside = hD >= 0 && hD < H && wD >= 0 && wD < W;
            if (inside) {
                const int v = disp[(size_t)hD * W + wD];
                if (v >= minD) {
                    nd = v;
                    const uint32_t qc = img4[(size_t)hD * W + wD];
                    nf = hsi ? hue_diff_u32(pc, qc) : color_diff_u32(pc, qc);
                    got = true;
                }
            }
        }
        occ_min = min(occ_min, nd);
        if (k == 0) { md = nd; mf = nf; }
        else if (mf < 0 || (mf > nf && nf > 0)) { md = nd; mf = nf; }
    }
    out[p] = (own == minD - 1) ? occ_min : md;  // the occlusion test is minD - 1 (:1209) although the marker written by the LRC is always -1
}

void proper_interpolation(const Launcher& L, const Dims& d, const int32_t* disp_in, int32_t* disp_out, const uint32_t* img4_left,
                          bool hsi)
{
    dim3 b(32, 8), g((d.W + 31) / 32, (d.H + 7) / 8);
    k_interpolate<<<g, b, 0, L.stream>>>(disp_in, disp_out, img4_left, d.H, d.W, hsi ? 1 : 0, d.minD);
    L.count(1);
}

// ========================================================================== a13
// discontinuityAdjustment (ADCensus.cpp:1241-1342) with integer restatements of the
// four OpenCV 4.13 calls (pinned against cv2 in tests/test_cvport.py):
//   gray = disp < 0 ? 0 : (uchar)disp   (wraps mod 256, :1249) -> equalizeHist (:1252)
//   -> blur 3x3 (:1263) -> Canny(30, 90, 3) (:1264) -> per-edge-pixel fix-up (:1268-1339).

__global__ void __launch_bounds__(256) k_gray_hist(const int32_t* __restrict__ disp, uint8_t* __restrict__ gray, int32_t* __restrict__ hist, size_t npx)
{
    __shared__ int sh[256];
    sh[threadIdx.x] = 0;
    __syncthreads();
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < npx; p += (size_t)gridDim.x * blockDim.x) {
        const int d = disp[p];
        const uint8_t g = d < 0 ? 0 : (uint8_t)d;
        gray[p] = g;
        atomicAdd(&sh[g], 1);
    }
    __syncthreads();
    if (sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], sh[threadIdx.x]);
}

// cv::equalizeHist LUT: scale = 255.f/(total - hist[first]); lut[j] = sat_u8(rint(sum * scale)).
__global__ void k_eq_lut(const int32_t* __restrict__ hist, int32_t* __restrict__ lut, int total)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int i = 0;
    while (i < 255 && !hist[i]) ++i;
    for (int j = 0; j < 256; ++j) lut[j] = 0;
    if (hist[i] == total) { lut[i] = i; return; }
    const float scale = __fdiv_rn(255.f, (float)(total - hist[i]));
    int sum = 0;
    for (lut[i++] = 0; i < 256; ++i) {
        sum += hist[i];
        int r = __float2int_rn(__fmul_rn((float)sum, scale));
        lut[i] = min(max(r, 0), 255);
    }
}

__device__ __forceinline__ int reflect101(int p, int n)
{
    if (p < 0) p = -p;
    if (p >= n) p = 2 * n - 2 - p;
    return p;
}

// cv::blur 3x3 of the equalised image: BORDER_REFLECT_101, round(sum/9) (never a tie).
__global__ void k_eq_blur(const uint8_t* __restrict__ gray, const int32_t* __restrict__ lut, uint8_t* __restrict__ blurred, int H, int W)
{
    __shared__ int slut[256];
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    if (tid < 256) slut[tid] = lut[tid];
    __syncthreads();
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    int s = 0;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
        const uint8_t* row = gray + (size_t)reflect101(y + dy, H) * W;
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) s += slut[row[reflect101(x + dx, W)]];
    }
    blurred[(size_t)y * W + x] = (uint8_t)((2 * s + 9) / 18);
}

// Sobel 3x3 (BORDER_REPLICATE) and L1 magnitude.
__global__ void k_sobel(const uint8_t* __restrict__ src, int16_t* __restrict__ gx, int16_t* __restrict__ gy, int32_t* __restrict__ mag, int H, int W)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    const int xm = max(x - 1, 0), xp = min(x + 1, W - 1), ym = max(y - 1, 0), yp = min(y + 1, H - 1);
    const uint8_t *r0 = src + (size_t)ym * W, *r1 = src + (size_t)y * W, *r2 = src + (size_t)yp * W;
    const int a = r0[xm], b = r0[x], c = r0[xp], d = r1[xm], f = r1[xp], g = r2[xm], h = r2[x], i = r2[xp];
    const int dx = (c - a) + 2 * (f - d) + (i - g);
    const int dy = (g - a) + 2 * (h - b) + (i - c);
    const size_t p = (size_t)y * W + x;
    gx[p] = (int16_t)dx;
    gy[p] = (int16_t)dy;
    mag[p] = abs(dx) + abs(dy);
}

// Non-maximum suppression + double threshold (imgproc/canny.cpp): 0 weak candidate, 1 none, 2 strong.
__global__ void k_canny_nms(const int16_t* __restrict__ gx, const int16_t* __restrict__ gy, const int32_t* __restrict__ mag,
                            uint8_t* __restrict__ map, int H, int W, int low, int high)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    auto M = [&](int yy, int xx) -> int { return (yy < 0 || yy >= H || xx < 0 || xx >= W) ? 0 : mag[(size_t)yy * W + xx]; };
    const size_t p = (size_t)y * W + x;
    const int m = mag[p];
    uint8_t v = 1;
    if (m > low) {
        const int xs = gx[p], ys = gy[p];
        const int ax = abs(xs), ay = abs(ys) << 15;
        const int tg22x = ax * 13573;
        bool is_max;
        if (ay < tg22x) {
            is_max = m > M(y, x - 1) && m >= M(y, x + 1);
        } else {
            const int tg67x = tg22x + (ax << 16);
            if (ay > tg67x) {
                is_max = m > M(y - 1, x) && m >= M(y + 1, x);
            } else {
                const int s = (xs ^ ys) < 0 ? -1 : 1;
                is_max = m > M(y - 1, x - s) && m > M(y + 1, x + s);
            }
        }
        if (is_max) v = m > high ? 2 : 0;
    }
    map[p] = v;
}

// Hysteresis: weak candidates 8-connected to a strong pixel become edges.  The result
// is order independent.  Cooperative persistent kernel: every CTA relaxes 32x32 tiles to
// a local fixed point in shared memory; grid-wide rounds repeat until no tile changed.
constexpr int HY_T = 32;
__global__ void __launch_bounds__(HY_T* HY_T) k_canny_hysteresis(uint8_t* __restrict__ map, int H, int W, int32_t* __restrict__ flags)
{
    cg::grid_group grid = cg::this_grid();
    __shared__ uint8_t t[HY_T + 2][HY_T + 2];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int tiles_x = (W + HY_T - 1) / HY_T, tiles_y = (H + HY_T - 1) / HY_T;
    const int ntiles = tiles_x * tiles_y;
    for (int round = 0;; ++round) {
        if (blockIdx.x == 0 && tx == 0 && ty == 0) flags[(round + 1) % 3] = 0;
        bool any_changed = false;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int x0 = (tile % tiles_x) * HY_T, y0 = (tile / tiles_x) * HY_T;
            for (int i = ty * HY_T + tx; i < (HY_T + 2) * (HY_T + 2); i += HY_T * HY_T) {
                const int ly = i / (HY_T + 2), lx = i % (HY_T + 2);
                const int y = y0 + ly - 1, x = x0 + lx - 1;
                t[ly][lx] = (y >= 0 && y < H && x >= 0 && x < W) ? map[(size_t)y * W + x] : 1;
            }
            __syncthreads();
            const bool inimg = (y0 + ty < H) && (x0 + tx < W);
            bool mine_changed = false;
            bool weak = inimg && t[ty + 1][tx + 1] == 0;
            if (__syncthreads_or(weak)) {
                for (;;) {
                    bool ch = false;
                    if (weak) {
                        const bool nb = t[ty][tx] == 2 || t[ty][tx + 1] == 2 || t[ty][tx + 2] == 2 || t[ty + 1][tx] == 2 ||
                                        t[ty + 1][tx + 2] == 2 || t[ty + 2][tx] == 2 || t[ty + 2][tx + 1] == 2 ||
                                        t[ty + 2][tx + 2] == 2;
                        if (nb) { ch = true; weak = false; mine_changed = true; }
                    }
                    __syncthreads();
                    if (ch) t[ty + 1][tx + 1] = 2;
                    if (!__syncthreads_or(ch)) break;
                }
                if (mine_changed) map[(size_t)(y0 + ty) * W + (x0 + tx)] = 2;
            }
            any_changed |= (bool)__syncthreads_or(mine_changed);
        }
        if (any_changed && tx == 0 && ty == 0) atomicExch(&flags[round % 3], 1);
        __threadfence();
        grid.sync();
        if (*((volatile int32_t*)&flags[round % 3]) == 0) break;
    }
}

__global__ void k_edges_finalize(const uint8_t* __restrict__ map, uint8_t* __restrict__ edges, size_t npx)
{
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < npx) edges[p] = map[p] == 2 ? 255 : 0;
}

__constant__ int c_adjH[8] = {-1, 1, -1, 1, -1, 1, 0, 0};
__constant__ int c_adjW[8] = {-1, 1, 0, 0, 1, -1, -1, 1};

__global__ void k_discont_adjust(const int32_t* __restrict__ disp, int32_t* __restrict__ out, const uint8_t* __restrict__ E,
                                 Vol vol, Dims dm)
{
    const int H = dm.H, W = dm.W;
    const int w = blockIdx.x * blockDim.x + threadIdx.x, h = blockIdx.y * blockDim.y + threadIdx.y;
    if (w >= W || h >= H) return;
    const size_t p = (size_t)h * W + w;
    int d = disp[p];
    if (h >= 1 && h < H - 1 && w >= 1 && w < W - 1 && E[p] != 0) {
        auto e = [&](int yy, int xx) -> bool { return E[(size_t)yy * W + xx] != 0; };
        int dir = -1;
        if (e(h - 1, w - 1) && e(h + 1, w + 1)) dir = 0;
        else if (e(h - 1, w + 1) && e(h + 1, w - 1)) dir = 4;
        else if (e(h - 1, w) || e(h + 1, w)) {
            if (e(h - 1, w - 1) || e(h - 1, w) || e(h - 1, w + 1))
                if (e(h + 1, w - 1) || e(h + 1, w) || e(h + 1, w + 1)) dir = 2;
        } else {
            if (e(h - 1, w - 1) || e(h, w - 1) || e(h + 1, w - 1))
                if (e(h - 1, w + 1) || e(h, w + 1) || e(h + 1, w + 1)) dir = 6;
        }
        const int m = dm.minD;  // valid: disp >= minD; the volume is read at plane disp - minD (:1310-1322)
        if (dir != -1 && d >= m) {
            dir = (dir + 4) % 8;
            float cost = *cell_ptr(vol, dm, p, d - m);
            const size_t p1 = (size_t)(h + c_adjH[dir]) * W + (w + c_adjW[dir]);
            const size_t p2 = (size_t)(h + c_adjH[dir + 1]) * W + (w + c_adjW[dir + 1]);
            const int d1 = disp[p1], d2 = disp[p2];
            const float c1 = d1 >= m ? *cell_ptr(vol, dm, p1, d1 - m) : -1.f;
            const float c2 = d2 >= m ? *cell_ptr(vol, dm, p2, d2 - m) : -1.f;
            if (c1 != -1.f && c1 < cost) { d = d1; cost = c1; }
            if (c2 != -1.f && c2 < cost) { d = d2; }
        }
    }
    out[p] = d;
}

cudaError_t discontinuity_adjustment(const Launcher& L, const Dims& d, const int32_t* disp_in, int32_t* disp_out,
                                     const Vol& vol_left, const EdgeScratch& s)
{
    const size_t npx = d.npx();
    dim3 b(32, 8), g((d.W + 31) / 32, (d.H + 7) / 8);
    cudaMemsetAsync(s.hist, 0, 256 * sizeof(int32_t), L.stream);
    k_gray_hist<<<592, 256, 0, L.stream>>>(disp_in, s.gray, s.hist, npx);
    k_eq_lut<<<1, 32, 0, L.stream>>>(s.hist, s.lut, (int)npx);
    k_eq_blur<<<g, b, 0, L.stream>>>(s.gray, s.lut, s.blurred, d.H, d.W);
    k_sobel<<<g, b, 0, L.stream>>>(s.blurred, s.gx, s.gy, s.mag, d.H, d.W);
    k_canny_nms<<<g, b, 0, L.stream>>>(s.gx, s.gy, s.mag, s.map, d.H, d.W, kCannyLow, kCannyHigh);
    L.count(5);
    {
        static PerDevice coop;
        if (coop.cur() == 0) {
            int dev = 0, sms = 0, per_sm = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_canny_hysteresis, HY_T * HY_T, 0);
            coop.cur() = (size_t)(sms * (per_sm > 0 ? 1 : 0));
            if (coop.cur() == 0) return cudaErrorLaunchOutOfResources;
        }
        const int coop_blocks = (int)coop.cur();
        cudaMemsetAsync(s.changed, 0, 3 * sizeof(int32_t), L.stream);
        uint8_t* map = s.map;
        int H = d.H, W = d.W;
        int32_t* flags = s.changed;
        void* args[] = {&map, &H, &W, &flags};
        cudaError_t e = cudaLaunchCooperativeKernel((void*)k_canny_hysteresis, dim3(coop_blocks), dim3(HY_T, HY_T), args, 0, L.stream);
        if (e != cudaSuccess) return e;
        L.count(1);
    }
    k_edges_finalize<<<(unsigned)((npx + 255) / 256), 256, 0, L.stream>>>(s.map, s.edges, npx);
    k_discont_adjust<<<g, b, 0, L.stream>>>(disp_in, disp_out, s.edges, vol_left, d);
    L.count(2);
    return cudaSuccess;
}

// ========================================================================== a14
// subpixelEnhancement (ADCensus.cpp:1344-1374): quadratic fit on the LEFT volume, then
// cv::medianBlur 3x3 (BORDER_REPLICATE) over the float map including negative markers.
__global__ void k_subpixel(const int32_t* __restrict__ disp, Vol vol, Dims dm, float* __restrict__ out)
{
    const size_t npx = dm.npx();
    const int Dn = dm.Dn;
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npx) return;
    const int d = disp[p];
    float f = (float)d;
    const int q = d - dm.minD;  // minD < d < maxD, planes d - minD and its neighbours (:1355-1362)
    if (q > 0 && q < Dn - 1) {
        const float cost = *cell_ptr(vol, dm, p, q), cp = *cell_ptr(vol, dm, p, q + 1), cm = *cell_ptr(vol, dm, p, q - 1);
        const float num = __fsub_rn(cp, cm);
        const float den = __fmul_rn(2.f, __fsub_rn(__fadd_rn(cp, cm), __fmul_rn(2.f, cost)));
        const float diff = __fdiv_rn(num, den);
        if (diff > -1.f && diff < 1.f) f = __fsub_rn(f, diff);
    }
    out[p] = f;
}

__device__ __forceinline__ void sort2(float& a, float& b)
{
    const float lo = fminf(a, b), hi = fmaxf(a, b);
    a = lo;
    b = hi;
}

__global__ void k_median3(const float* __restrict__ src, float* __restrict__ dst, int H, int W)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    float p[9];
    int k = 0;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
        const float* row = src + (size_t)min(max(y + dy, 0), H - 1) * W;
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) p[k++] = row[min(max(x + dx, 0), W - 1)];
    }
    sort2(p[1], p[2]); sort2(p[4], p[5]); sort2(p[7], p[8]);
    sort2(p[0], p[1]); sort2(p[3], p[4]); sort2(p[6], p[7]);
    sort2(p[1], p[2]); sort2(p[4], p[5]); sort2(p[7], p[8]);
    sort2(p[0], p[3]); sort2(p[5], p[8]); sort2(p[4], p[7]);
    sort2(p[3], p[6]); sort2(p[1], p[4]); sort2(p[2], p[5]);
    sort2(p[4], p[7]); sort2(p[4], p[2]); sort2(p[6], p[4]);
    sort2(p[4], p[2]);
    dst[(size_t)y * W + x] = p[4];
}

void subpixel(const Launcher& L, const Dims& d, const int32_t* disp, const Vol& vol_left, float* tmp, float* out)
{
    const size_t npx = d.npx();
    k_subpixel<<<(unsigned)((npx + 255) / 256), 256, 0, L.stream>>>(disp, vol_left, d, tmp);
    dim3 b(32, 8), g((d.W + 31) / 32, (d.H + 7) / 8);
    k_median3<<<g, b, 0, L.stream>>>(tmp, out, d.H, d.W);
    L.count(2);
}

}  // namespace tsm
