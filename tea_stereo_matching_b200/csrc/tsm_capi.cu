// tsm_capi.cu -- the C-ABI of include/tsm.h: context, device arena, stage
// orchestration.  Mirrors the call sequence of stereo::ADCensus::compute
// (reference source/ADCensus.cpp:330-407) and multiOptimize (:1376-1392).
#include "tsm_common.cuh"
#include <algorithm>
#include <charconv>
#include <limits>
#include <memory>
#include <mutex>
#include <thread>
#include <limits.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>

using namespace tsm;

namespace {
thread_local std::string g_create_error;

struct Buf {
    void* p = nullptr;
    size_t bytes = 0;
};

struct StageTimer {
    const char* name;
    cudaEvent_t beg, end;
};
}  // namespace

struct tsm_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t side = nullptr;        // second stream for work independent of the main chain (Launcher::side)
    cudaEvent_t fork_join[2] = {nullptr, nullptr};
    std::string err;
    long long launches = 0;

    // geometry of the arena
    Dims dm{0, 0, 0, 0, 0, 0};
    tsm_adcensus_config cfg{};
    bool have_pair = false;

    // device buffers
    Buf img[2], img4[2], census[2], arms[2], desc_h[2], desc_v[2], fdesc_h[2], fdesc_v[2], flags[2], tflags[2], sbits[2], vol[2], vtail[2], wta_[2];
    Buf dense;  // [H][W][Dn] staging for volume taps / pokes
    bool stage_mode = false;  // tsm_stage_run: keep every tap-able buffer complete
    Buf disp[2], fin, ftmp;
    Buf v_vote, v_lowcnt, v_off, v_start, v_sums, v_flat, v_stash, v_pre;
    Buf e_gray, e_blur, e_mag, e_gx, e_gy, e_map, e_edges, e_hist, e_lut, e_changed;
    Buf tab_ad, tab_c, agg_ctr, tab_ad_hsi, hsi_lut;
    bool use_scan3 = false;  // the flags of the blocked scanline walk (k_scanline3.cu) were prepared for the current pair
    bool mask = false;     // mask matching mode: black pixels are holes (cost 2, zero arms, skipped scanline steps)
    bool roi = false;      // ROI / mask matching mode: maxD = W / 2, HSI hue filter instead of the Gauss-median, offset + final marking
    int roi_offset = 0;
    bool hsi = false, hsi_ready = false;  // colour model of the current configuration; HSI tables uploaded
    Buf k_in, k_out, k_tab, k_range;  // disparity consumers: staged input map, output, colour table, min/max
    int fin_H = 0, fin_W = 0;        // geometry of the map in `fin` (0 = none yet)
    int disp_cur = 0;  // which of disp[2] holds the working map
    float p1_lo = 0.f, p2_lo = 0.f;
    bool tables_ready = false;

    // pinned staging
    uint8_t* h_pair = nullptr;  // left | right packed
    float* h_out = nullptr;
    size_t h_pair_bytes = 0, h_out_bytes = 0;
    bool pending = false;

    // rectify
    Buf r_src, r_map1[2], r_map2[2], r_fmap[2][2];
    // device copies of rectify maps, one slot per map pair; valid while the caller's generation id is unchanged
    struct MapSlot {
        unsigned long long gen = 0;  // 0 = empty
        const void* m1 = nullptr;
        const void* m2 = nullptr;
        int kind = -1, H = 0, W = 0;
    } map_slot[2];
    int remap_victim = 0;
    uint8_t* h_stereo = nullptr;
    size_t h_stereo_bytes = 0;

    // profiling
    bool profiling = false;
    std::vector<StageTimer> timers;
    size_t timers_used = 0;
    size_t sub_open = (size_t)-1;  // timer of the launch-level interval that is open (Launcher::begin / end)
};

namespace {

int fail(tsm_ctx* c, int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf;
    else g_create_error = buf;
    return code;
}

#define CK(c, call)                                                                                      \
    do {                                                                                                 \
        cudaError_t e__ = (call);                                                                        \
        if (e__ != cudaSuccess)                                                                          \
            return fail((c), e__ == cudaErrorMemoryAllocation ? TSM_E_OOM : TSM_E_CUDA, "%s: %s", #call, \
                        cudaGetErrorString(e__));                                                        \
    } while (0)

int ensure(tsm_ctx* c, Buf& b, size_t bytes, bool zero = false)
{
    if (b.bytes >= bytes && b.p) return TSM_OK;
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.bytes = 0;
    cudaError_t e = cudaMalloc(&b.p, bytes);
    if (e != cudaSuccess) return fail(c, TSM_E_OOM, "cudaMalloc(%zu bytes): %s", bytes, cudaGetErrorString(e));
    b.bytes = bytes;
    if (zero) {
        e = cudaMemsetAsync(b.p, 0, bytes, c->stream);
        if (e != cudaSuccess) return fail(c, TSM_E_CUDA, "cudaMemsetAsync: %s", cudaGetErrorString(e));
    }
    return TSM_OK;
}
void release(Buf& b)
{
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.bytes = 0;
}

int ensure_pinned(tsm_ctx* c, void** p, size_t* have, size_t bytes)
{
    if (*p && *have >= bytes) return TSM_OK;
    if (*p) cudaFreeHost(*p);
    *p = nullptr;
    *have = 0;
    cudaError_t e = cudaMallocHost(p, bytes);
    if (e != cudaSuccess) return fail(c, TSM_E_OOM, "cudaMallocHost(%zu bytes): %s", bytes, cudaGetErrorString(e));
    *have = bytes;
    return TSM_OK;
}

int check_cfg(tsm_ctx* c, const tsm_adcensus_config* cfg, int H, int W)
{
    if (!cfg) return fail(c, TSM_E_ARG, "[ADCensus] config is NULL");
    // setMinMaxDisparity, ADCensus.cpp:309-310
    if ((long long)cfg->min_disparity * cfg->max_disparity < 0 || cfg->min_disparity >= cfg->max_disparity)
        return fail(c, TSM_E_ARG, "[ADCensus] Set MinMaxDisparity error.");
    if (cfg->offset < 0) return fail(c, TSM_E_ARG, "[ADCensus] Offset must be positive.");  // ADCensus.cpp:325-326
    if (cfg->min_disparity < 0)
        return fail(c, TSM_E_UNSUPPORTED, "[ADCensus] negative disparities index the reference's cost planes out of range (ADCensus.cpp:1398)");
    // cost2disparity walks the planes minD .. maxD - minD (ADCensus.cpp:1398): with maxD < 2 minD the loop is empty and the
    // reference returns uninitialised memory
    if (!(cfg->roi_matching || cfg->mask_matching) && cfg->max_disparity < 2 * (long long)cfg->min_disparity)
        return fail(c, TSM_E_UNSUPPORTED, "[ADCensus] max_disparity < 2 * min_disparity: the reference's WTA range is empty");
    // ROI mode searches the whole half width: maxDisparity = width / 2 at compute time (ADCensus.cpp:339-340)
    if ((cfg->roi_matching || cfg->mask_matching) && W / 2 < 2 * (long long)cfg->min_disparity)
        return fail(c, TSM_E_UNSUPPORTED, "[ADCensus] W / 2 < 2 * min_disparity: the reference's WTA range is empty");
    if (((cfg->roi_matching || cfg->mask_matching) ? W / 2 : cfg->max_disparity) - cfg->min_disparity + 1 > kMaxLevels)
        return fail(c, TSM_E_UNSUPPORTED, "[ADCensus] more than %d disparity levels", kMaxLevels);
    if ((cfg->roi_matching || cfg->mask_matching) && W / 2 < 1) return fail(c, TSM_E_ARG, "[ADCensus] Image error.");
    if (H <= 0 || W <= 0) return fail(c, TSM_E_ARG, "[ADCensus] Image error.");  // ADCensus.cpp:332-333
    if (cfg->max_disparity > 65535) return fail(c, TSM_E_ARG, "[ADCensus] max_disparity too large");
    return TSM_OK;
}

// bgr2hsi (ADCensus.cpp:1429-1473) for every 24-bit pixel, evaluated ONCE per process with the host's libm in the
// reference's expression order (float arithmetic, the hue division in double because CV_PI is a double literal,
// float -> uchar conversions truncate).  Index = B | G << 8 | R << 16, value = H | S << 8 | I << 16.
static const std::vector<uint32_t>& hsi_table()
{
    static std::vector<uint32_t> tab;
    static std::once_flag once;
    std::call_once(once, [] {
        tab.resize(1u << 24);
        const double two_pi = 2 * 3.1415926535897932384626433832795;  // 2 * CV_PI
        auto fill = [&](uint32_t lo, uint32_t hi) {
        for (uint32_t i = lo; i < hi; ++i) {
            const volatile float blueValue = (float)(i & 0xff) / 255.f, greenValue = (float)((i >> 8) & 0xff) / 255.f,
                                 redValue = (float)((i >> 16) & 0xff) / 255.f;
            const volatile float sum0 = blueValue + greenValue;
            const volatile float sum = sum0 + redValue;
            const volatile float iValue = sum / 3.0f;
            float sValue, hValue;
            if (sum == 0) sValue = 0;
            else {
                const float mn = std::min(std::min((float)blueValue, (float)greenValue), (float)redValue);
                const volatile float t3 = 3 * mn;
                const volatile float q = t3 / sum;
                sValue = 1 - q;
            }
            const volatile float rg = redValue - greenValue, rb = redValue - blueValue, gb = greenValue - blueValue;
            const volatile float a2 = rg * rg, b2 = rb * gb;
            const volatile float rad = a2 + b2;
            const float den = sqrtf(rad);
            const volatile float n0 = 2 * redValue;
            const volatile float n1 = n0 - greenValue;
            const volatile float n2 = n1 - blueValue;
            const float num = n2 / 2.f;
            if (den == 0.f || den <= num || sValue < 0.05f) hValue = 0;
            else {
                const volatile float ratio = num / den;
                const float theta = acosf(ratio);
                hValue = blueValue <= greenValue ? (float)(theta / two_pi) : (float)(1 - theta / two_pi);
            }
            const volatile float hi = hValue * 255, si = sValue * 255, ii = iValue * 255;
            const uint32_t H = (unsigned char)hi, S = (unsigned char)si, I = (unsigned char)ii;
            tab[i] = H | (S << 8) | (I << 16);
        }
        };
        const unsigned nt = std::max(1u, std::min(8u, std::thread::hardware_concurrency()));
        std::vector<std::thread> pool;
        const uint32_t per = (1u << 24) / nt;
        for (unsigned t = 0; t < nt; ++t) pool.emplace_back(fill, t * per, t + 1 == nt ? (1u << 24) : (t + 1) * per);
        for (auto& th : pool) th.join();
    });
    return tab;
}

int ensure_hsi_tables(tsm_ctx* c)
{
    if (c->hsi_ready) return TSM_OK;
    // exp(-adCost / lambdaAD) over t = 2 * adCost (adCost is a multiple of 0.5, computeHSIADCost :439-452)
    std::vector<float> tab(kTabAdHsi);
    const volatile float lambda_ad = 10.f, half = 0.5f;
    for (int t = 0; t < kTabAdHsi; ++t) {
        volatile float ad = (float)t * half;
        tab[t] = expf(-ad / lambda_ad);
    }
    const std::vector<uint32_t>& lut = hsi_table();
    int rc;
    if ((rc = ensure(c, c->tab_ad_hsi, tab.size() * 4))) return rc;
    if ((rc = ensure(c, c->hsi_lut, lut.size() * 4))) return rc;
    CK(c, cudaMemcpyAsync(c->tab_ad_hsi.p, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice, c->stream));
    CK(c, cudaMemcpyAsync(c->hsi_lut.p, lut.data(), lut.size() * 4, cudaMemcpyHostToDevice, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    c->hsi_ready = true;
    return TSM_OK;
}

// Host-side LUTs with the host's own expf, in the reference's expression order:
//   expf(-((float)s / 3.f) / 10.f)  (ADCensus.cpp:435, :518)  and  expf(-(float)n / 30.f).
int ensure_tables(tsm_ctx* c)
{
    if (c->tables_ready) return TSM_OK;
    float tab_ad[766], tab_c[192];
    const volatile float lambda_ad = 10.f, lambda_c = 30.f, three = 3.f;
    for (int s = 0; s < 766; ++s) {
        volatile float ad = (float)s / three;
        tab_ad[s] = expf(-ad / lambda_ad);
    }
    for (int n = 0; n < 192; ++n) {
        volatile float cn = (float)n;
        tab_c[n] = expf(-cn / lambda_c);
    }
    const volatile float pi1 = 1.f, pi2 = 3.f, ten = 10.f;
    c->p1_lo = pi1 / ten;  // ADCensus.cpp:976-977
    c->p2_lo = pi2 / ten;
    int rc;
    if ((rc = ensure(c, c->tab_ad, sizeof tab_ad))) return rc;
    if ((rc = ensure(c, c->tab_c, sizeof tab_c))) return rc;
    CK(c, cudaMemcpyAsync(c->tab_ad.p, tab_ad, sizeof tab_ad, cudaMemcpyHostToDevice, c->stream));
    CK(c, cudaMemcpyAsync(c->tab_c.p, tab_c, sizeof tab_c, cudaMemcpyHostToDevice, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));  // the host arrays are on this stack frame
    c->tables_ready = true;
    return TSM_OK;
}

int ensure_arena(tsm_ctx* c, const tsm_adcensus_config* cfg, int H, int W)
{
    int rc = check_cfg(c, cfg, H, W);
    if (rc) return rc;
    CK(c, cudaSetDevice(c->device));
    if ((rc = ensure_tables(c))) return rc;
    c->hsi = cfg->color_model == TSM_COLOR_HSI;
    if (c->hsi && (rc = ensure_hsi_tables(c))) return rc;
    Dims d;
    c->roi = cfg->roi_matching != 0 || cfg->mask_matching != 0;  // both: maxD = W / 2, hue filter, offset, final marking
    c->mask = cfg->mask_matching != 0;
    c->roi_offset = cfg->offset;
    d.set(H, W, (c->roi ? W / 2 : cfg->max_disparity) - cfg->min_disparity + 1, cfg->min_disparity);
    c->dm = d;
    c->cfg = *cfg;
    const size_t npx = d.npx();
    for (int k = 0; k < 2; ++k) {
        if ((rc = ensure(c, c->img[k], npx * 3))) return rc;
        if ((rc = ensure(c, c->img4[k], npx * 4))) return rc;
        if ((rc = ensure(c, c->census[k], npx * 6 * 8))) return rc;
        if ((rc = ensure(c, c->arms[k], npx * 4))) return rc;
        if ((rc = ensure(c, c->desc_h[k], d.desc_h_words() * 8, true))) return rc;
        if ((rc = ensure(c, c->desc_v[k], d.desc_v_words() * 8, true))) return rc;
        if ((rc = ensure(c, c->fdesc_h[k], d.fdesc_h_words() * 4, true))) return rc;
        if ((rc = ensure(c, c->fdesc_v[k], d.fdesc_v_words() * 4, true))) return rc;
        if ((rc = ensure(c, c->flags[k], npx))) return rc;
        if ((rc = ensure(c, c->tflags[k], ((size_t)2 * H * d.stab_pitch() + 64) * 4, true))) return rc;
        if (scanline3_geometry(d) && (rc = ensure(c, c->sbits[k], (sb_layout(H, W).words + 64) * 4, true))) return rc;
        if ((rc = ensure(c, c->vol[k], (npx * d.Dm + aggregate_overread_floats(d)) * 4 + 256, true))) return rc;
        if ((rc = ensure(c, c->vtail[k], (npx * d.Rp + aggregate_overread_floats(d)) * 4 + 256, true))) return rc;
        if ((rc = ensure(c, c->wta_[k], npx * 4))) return rc;
        if ((rc = ensure(c, c->disp[k], npx * 4))) return rc;
    }
    if ((rc = ensure(c, c->agg_ctr, kAggCounterBytes))) return rc;
    if ((rc = ensure(c, c->fin, npx * 4))) return rc;
    if ((rc = ensure(c, c->ftmp, npx * 4))) return rc;
    if ((rc = ensure(c, c->v_vote, npx * 4))) return rc;
    if ((rc = ensure(c, c->v_lowcnt, npx * 4))) return rc;
    if ((rc = ensure(c, c->v_off, npx * 4))) return rc;
    if ((rc = ensure(c, c->v_start, npx * 4))) return rc;
    if ((rc = ensure(c, c->v_pre, npx + 256))) return rc;
    if ((rc = ensure(c, c->v_sums, (npx / 2048 + 2) * 8))) return rc;
    if ((rc = ensure(c, c->v_flat, npx * kVotingThresh * 2))) return rc;
    if ((rc = ensure(c, c->v_stash, npx * kVotingThresh * 2))) return rc;
    if ((rc = ensure(c, c->e_gray, npx))) return rc;
    if ((rc = ensure(c, c->e_blur, npx))) return rc;
    if ((rc = ensure(c, c->e_mag, npx * 4))) return rc;
    if ((rc = ensure(c, c->e_gx, npx * 2))) return rc;
    if ((rc = ensure(c, c->e_gy, npx * 2))) return rc;
    if ((rc = ensure(c, c->e_map, npx))) return rc;
    if ((rc = ensure(c, c->e_edges, npx))) return rc;
    if ((rc = ensure(c, c->e_hist, 256 * 4))) return rc;
    if ((rc = ensure(c, c->e_lut, 256 * 4))) return rc;
    if ((rc = ensure(c, c->e_changed, 4 * 4))) return rc;
    return TSM_OK;
}

ViewPtrs view_ptrs(tsm_ctx* c, int k)
{
    ViewPtrs v;
    v.img = (const uint8_t*)c->img[k].p;
    v.img4 = (const uint32_t*)c->img4[k].p;
    v.census = (const uint64_t*)c->census[k].p;
    v.arms = (const uchar4*)c->arms[k].p;
    v.desc_h = (const uint32_t*)c->desc_h[k].p;
    v.desc_v = (const uint32_t*)c->desc_v[k].p;
    v.rcp_h = (const float*)(v.desc_h + c->dm.desc_h_words());
    v.rcp_v = (const float*)(v.desc_v + c->dm.desc_v_words());
    v.fdesc_h = (const uint32_t*)c->fdesc_h[k].p + kFdescFront;
    v.fdesc_v = (const uint32_t*)c->fdesc_v[k].p + kFdescFront;
    v.flags = (const uint8_t*)c->flags[k].p;
    v.stab = (const uint32_t*)c->tflags[k].p;
    v.sbits = (const uint32_t*)c->sbits[k].p;
    v.vol.main = (float*)c->vol[k].p;
    v.vol.tail = (float*)c->vtail[k].p;
    return v;
}

size_t timer_open(tsm_ctx* c, const char* name);

struct ScopedStage {
    tsm_ctx* c;
    size_t idx = (size_t)-1;
    ScopedStage(tsm_ctx* ctx, const char* name) : c(ctx)
    {
        if (!c->profiling) return;
        idx = timer_open(c, name);
    }
    ~ScopedStage()
    {
        if (idx != (size_t)-1) cudaEventRecord(c->timers[idx].end, c->stream);
    }
};

size_t timer_open(tsm_ctx* c, const char* name)
{
    if (c->timers_used == c->timers.size()) {
        StageTimer t;
        t.name = name;
        cudaEventCreate(&t.beg);
        cudaEventCreate(&t.end);
        c->timers.push_back(t);
    }
    const size_t idx = c->timers_used++;
    c->timers[idx].name = name;
    cudaEventRecord(c->timers[idx].beg, c->stream);
    return idx;
}

// Launcher hook: one event pair around a single launch (or a few) inside a stage
void prof_mark(void* self, const char* name, int begin)
{
    tsm_ctx* c = (tsm_ctx*)self;
    if (!c->profiling) return;
    if (begin) c->sub_open = timer_open(c, name);
    else if (c->sub_open != (size_t)-1) {
        cudaEventRecord(c->timers[c->sub_open].end, c->stream);
        c->sub_open = (size_t)-1;
    }
}

Launcher make_launcher(tsm_ctx* c)
{
    Launcher L{c->stream, &c->launches};
    if (c->profiling) { L.prof = c; L.mark = prof_mark; }
    if (c->side && c->fork_join[0] && c->fork_join[1]) { L.side = c->side; L.events = c->fork_join; }
    return L;
}

// Runs the stages in `mask` on the current arena, in pipeline order.
int run_stages(tsm_ctx* c, int mask, int arg)
{
    const Dims& d = c->dm;
    Launcher L = make_launcher(c);
    ViewPtrs vl = view_ptrs(c, 0), vr = view_ptrs(c, 1);
    if (c->profiling && mask == TSM_STAGE_ALL && arg != -2) c->timers_used = 0;  // -2: the caller already recorded a stage
    if (mask & TSM_STAGE_PREP) {
        ScopedStage s(c, "prep");
        const uint8_t* img[2]; uint32_t* img4[2]; uint64_t* census[2]; uchar4* arms[2];
        uint32_t *desc_h[2], *desc_v[2], *fdesc_h[2], *fdesc_v[2]; uint8_t* flags[2];
        for (int k = 0; k < 2; ++k) {
            img[k] = (const uint8_t*)c->img[k].p; img4[k] = (uint32_t*)c->img4[k].p; census[k] = (uint64_t*)c->census[k].p;
            arms[k] = (uchar4*)c->arms[k].p; desc_h[k] = (uint32_t*)c->desc_h[k].p; desc_v[k] = (uint32_t*)c->desc_v[k].p;
            fdesc_h[k] = (uint32_t*)c->fdesc_h[k].p; fdesc_v[k] = (uint32_t*)c->fdesc_v[k].p; flags[k] = (uint8_t*)c->flags[k].p;
        }
        prep_views(L, d, img, img4, census, arms, desc_h, desc_v, fdesc_h, fdesc_v, flags, model_params(c->hsi, c->mask),
                   (const uint32_t*)c->hsi_lut.p, c->roi);
        c->use_scan3 = scanline3_supported(d);  // the scanline stage follows what was prepared
        if (c->use_scan3)
            prep_scan_bits(L, d, (const uint8_t*)c->flags[0].p, (const uint8_t*)c->flags[1].p, (uint32_t*)c->sbits[0].p,
                           (uint32_t*)c->sbits[1].p);
        else
            prep_scan_tables(L, d, (const uint8_t*)c->flags[0].p, (const uint8_t*)c->flags[1].p, (uint32_t*)c->tflags[0].p,
                             (uint32_t*)c->tflags[1].p);
    }
    if (mask & TSM_STAGE_INIT) {
        ScopedStage s(c, "cost_init");
        if (d.minD != 0)
            cost_init_general(L, d, vl, vr, (const float*)(c->hsi ? c->tab_ad_hsi.p : c->tab_ad.p), (const float*)c->tab_c.p, c->hsi, c->mask);
        else
            cost_init(L, d, vl, vr, (const float*)(c->hsi ? c->tab_ad_hsi.p : c->tab_ad.p), (const float*)c->tab_c.p, c->hsi, c->mask);
    }
    if (mask & TSM_STAGE_AGGREGATE) {
        ScopedStage s(c, "aggregate");
        aggregate(L, d, vl, vr, (unsigned*)c->agg_ctr.p);
    }
    if (mask & TSM_STAGE_SCANLINE) {
        ScopedStage s(c, "scanline");
        // the last (leftward) pass also writes both WTA maps (cost2disparity fused)
        // (minD != 0: the WTA range is restricted, the stand-alone kernel below does it and needs the right volume's last store)
        if (c->use_scan3)
            scanline3(L, d, vl, vr, c->p1_lo, c->p2_lo, (int32_t*)c->wta_[0].p, (int32_t*)c->wta_[1].p, c->stage_mode || d.minD != 0);
        else
            scanline(L, d, vl, vr, c->p1_lo, c->p2_lo, (int32_t*)c->wta_[0].p, (int32_t*)c->wta_[1].p, c->stage_mode || d.minD != 0);
    }
    if (((mask & TSM_STAGE_WTA) && !(mask & TSM_STAGE_SCANLINE)) || ((mask & TSM_STAGE_SCANLINE) && d.minD != 0)) {
        ScopedStage s(c, "wta");
        wta(L, d, vl.vol, (int32_t*)c->wta_[0].p);
        wta(L, d, vr.vol, (int32_t*)c->wta_[1].p);
    }
    if (mask & TSM_STAGE_LRC) {
        ScopedStage s(c, "lrc");
        c->disp_cur = 0;
        lrc(L, d, (const int32_t*)c->wta_[0].p, (const int32_t*)c->wta_[1].p, (int32_t*)c->disp[0].p);
    }
    if (mask & TSM_STAGE_VOTE) {
        ScopedStage s(c, "region_voting");
        VoteScratch vs;
        vs.vote = (int32_t*)c->v_vote.p;
        vs.lowcnt = (int32_t*)c->v_lowcnt.p;
        vs.off = (int32_t*)c->v_off.p;
        vs.start = (int32_t*)c->v_start.p;
        vs.pre = (uint8_t*)c->v_pre.p;
        vs.blocksums = (int32_t*)c->v_sums.p;
        vs.flat = (uint16_t*)c->v_flat.p;
        vs.stash = (uint16_t*)c->v_stash.p;
        vs.flat_capacity = c->v_flat.bytes / 2;
        // multiOptimize: 5 calls, horizontalFirst = F,T,F,T,F (ADCensus.cpp:1382-1387)
        for (int i = 0; i < 5; ++i) {
            if (arg >= 0 && arg != i) continue;
            const bool hf = (i & 1) != 0;
            region_voting(L, d, (const int32_t*)c->disp[c->disp_cur].p, (int32_t*)c->disp[c->disp_cur ^ 1].p, vl.arms, hf, vs);
            c->disp_cur ^= 1;
        }
    }
    if (mask & TSM_STAGE_INTERP) {
        ScopedStage s(c, "interpolation");
        proper_interpolation(L, d, (const int32_t*)c->disp[c->disp_cur].p, (int32_t*)c->disp[c->disp_cur ^ 1].p, vl.img4, c->hsi);
        c->disp_cur ^= 1;
    }
    if (mask & TSM_STAGE_DISCONT) {
        ScopedStage s(c, "discontinuity");
        EdgeScratch es;
        es.gray = (uint8_t*)c->e_gray.p;
        es.blurred = (uint8_t*)c->e_blur.p;
        es.mag = (int32_t*)c->e_mag.p;
        es.gx = (int16_t*)c->e_gx.p;
        es.gy = (int16_t*)c->e_gy.p;
        es.map = (uint8_t*)c->e_map.p;
        es.edges = (uint8_t*)c->e_edges.p;
        es.hist = (int32_t*)c->e_hist.p;
        es.lut = (int32_t*)c->e_lut.p;
        es.changed = (int32_t*)c->e_changed.p;
        es.h_changed = nullptr;
        cudaError_t e = discontinuity_adjustment(L, d, (const int32_t*)c->disp[c->disp_cur].p,
                                                 (int32_t*)c->disp[c->disp_cur ^ 1].p, vl.vol, es);
        if (e != cudaSuccess) return fail(c, TSM_E_CUDA, "discontinuity_adjustment: %s", cudaGetErrorString(e));
        c->disp_cur ^= 1;
    }
    if (mask & TSM_STAGE_SUBPIXEL) {
        ScopedStage s(c, "subpixel");
        subpixel(L, d, (const int32_t*)c->disp[c->disp_cur].p, vl.vol, (float*)c->ftmp.p, (float*)c->fin.p);
        if (c->roi) roi_finish(L, d, (float*)c->fin.p, (const uint8_t*)c->img[0].p, c->roi_offset);
        c->fin_H = d.H;
        c->fin_W = d.W;
    }
    CK(c, cudaGetLastError());
    return TSM_OK;
}

int upload_pair_host(tsm_ctx* c, const uint8_t* left, size_t lstep, const uint8_t* right, size_t rstep, int H, int W)
{
    const size_t row = (size_t)W * 3, img_bytes = row * H;
    int rc = ensure_pinned(c, (void**)&c->h_pair, &c->h_pair_bytes, 2 * img_bytes);
    if (rc) return rc;
    for (int y = 0; y < H; ++y) {
        memcpy(c->h_pair + (size_t)y * row, left + (size_t)y * lstep, row);
        memcpy(c->h_pair + img_bytes + (size_t)y * row, right + (size_t)y * rstep, row);
    }
    CK(c, cudaMemcpyAsync(c->img[0].p, c->h_pair, img_bytes, cudaMemcpyHostToDevice, c->stream));
    CK(c, cudaMemcpyAsync(c->img[1].p, c->h_pair + img_bytes, img_bytes, cudaMemcpyHostToDevice, c->stream));
    return TSM_OK;
}

int check_images(tsm_ctx* c, const void* left, size_t lstep, const void* right, size_t rstep, int H, int W)
{
    if (!left || !right || H <= 0 || W <= 0) return fail(c, TSM_E_ARG, "[ADCensus] Image error.");
    if (lstep < (size_t)W * 3 || rstep < (size_t)W * 3) return fail(c, TSM_E_ARG, "[ADCensus] Image error (row stride < 3*W).");
    return TSM_OK;
}

}  // namespace

extern "C" {

int tsm_version(void) { return TSM_VERSION; }

const char* tsm_status_string(int s)
{
    switch (s) {
        case TSM_OK: return "ok";
        case TSM_E_ARG: return "invalid argument";
        case TSM_E_CUDA: return "CUDA error";
        case TSM_E_OOM: return "out of memory";
        case TSM_E_UNSUPPORTED: return "unsupported configuration";
        case TSM_E_STATE: return "invalid call sequence";
        default: return "unknown status";
    }
}

const char* tsm_last_error(const tsm_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int tsm_device_count(int* count)
{
    if (!count) return TSM_E_ARG;
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) {
        *count = 0;
        return fail(nullptr, TSM_E_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    }
    return TSM_OK;
}

int tsm_create_on_stream(int device, void* cuda_stream, tsm_ctx** out)
{
    if (!out) return fail(nullptr, TSM_E_ARG, "tsm_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(nullptr, TSM_E_CUDA, "tsm_create: no CUDA device (%s); this library has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= n) return fail(nullptr, TSM_E_ARG, "tsm_create: device %d out of range [0,%d)", device, n);
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail(nullptr, TSM_E_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    tsm_ctx* c = new tsm_ctx();
    c->device = device;
    if (cuda_stream) {
        c->stream = (cudaStream_t)cuda_stream;
    } else {
        e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) {
            delete c;
            return fail(nullptr, TSM_E_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
        }
        c->own_stream = true;
    }
    // side stream + fork / join events (best effort: without them everything stays on the one stream)
    if (cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking) != cudaSuccess) c->side = nullptr;
    for (auto& e : c->fork_join)
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) e = nullptr;
    *out = c;
    return TSM_OK;
}

int tsm_create(int device, tsm_ctx** out) { return tsm_create_on_stream(device, nullptr, out); }

void tsm_destroy(tsm_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    Buf* all[] = {&c->img[0], &c->img[1], &c->img4[0], &c->img4[1], &c->census[0], &c->census[1], &c->arms[0], &c->arms[1],
                  &c->desc_h[0], &c->desc_h[1], &c->desc_v[0], &c->desc_v[1], &c->fdesc_h[0], &c->fdesc_h[1], &c->fdesc_v[0], &c->fdesc_v[1], &c->flags[0], &c->flags[1], &c->tflags[0], &c->tflags[1], &c->sbits[0], &c->sbits[1], &c->vtail[0], &c->vtail[1], &c->dense, &c->vol[0], &c->vol[1], &c->wta_[0], &c->wta_[1],
                  &c->disp[0], &c->disp[1], &c->fin, &c->ftmp, &c->v_vote, &c->v_lowcnt, &c->v_off, &c->v_start, &c->v_pre, &c->v_stash,
                  &c->v_sums, &c->v_flat, &c->e_gray, &c->e_blur, &c->e_mag, &c->e_gx, &c->e_gy, &c->e_map, &c->e_edges,
                  &c->e_hist, &c->e_lut, &c->e_changed, &c->tab_ad, &c->tab_c, &c->agg_ctr, &c->tab_ad_hsi, &c->hsi_lut, &c->k_in, &c->k_out, &c->k_tab, &c->k_range, &c->r_src, &c->r_map1[0], &c->r_map1[1],
                  &c->r_map2[0], &c->r_map2[1], &c->r_fmap[0][0], &c->r_fmap[0][1], &c->r_fmap[1][0], &c->r_fmap[1][1]};
    for (Buf* b : all) release(*b);
    if (c->h_pair) cudaFreeHost(c->h_pair);
    if (c->h_out) cudaFreeHost(c->h_out);
    if (c->h_stereo) cudaFreeHost(c->h_stereo);
    for (auto& t : c->timers) {
        cudaEventDestroy(t.beg);
        cudaEventDestroy(t.end);
    }
    if (c->side) { cudaStreamSynchronize(c->side); cudaStreamDestroy(c->side); }
    for (auto& e : c->fork_join)
        if (e) cudaEventDestroy(e);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
}

int tsm_synchronize(tsm_ctx* c)
{
    if (!c) return TSM_E_ARG;
    CK(c, cudaStreamSynchronize(c->stream));
    return TSM_OK;
}

int tsm_adcensus_compute_device(tsm_ctx* c, const tsm_adcensus_config* cfg, const uint8_t* d_left, const uint8_t* d_right,
                                int H, int W, float* d_disparity)
{
    if (!c) return TSM_E_ARG;
    if (!d_left || !d_right || !d_disparity) return fail(c, TSM_E_ARG, "[ADCensus] Image error.");
    if (c->pending) return fail(c, TSM_E_STATE, "tsm_adcensus_compute_device: a pair is enqueued on this context and not waited for");
    int rc = ensure_arena(c, cfg, H, W);
    if (rc) return rc;
    const size_t img_bytes = (size_t)H * W * 3;
    CK(c, cudaMemcpyAsync(c->img[0].p, d_left, img_bytes, cudaMemcpyDeviceToDevice, c->stream));
    CK(c, cudaMemcpyAsync(c->img[1].p, d_right, img_bytes, cudaMemcpyDeviceToDevice, c->stream));
    c->have_pair = true;
    if ((rc = run_stages(c, TSM_STAGE_ALL, -1))) return rc;
    CK(c, cudaMemcpyAsync(d_disparity, c->fin.p, (size_t)H * W * 4, cudaMemcpyDeviceToDevice, c->stream));
    return TSM_OK;
}

int tsm_adcensus_enqueue(tsm_ctx* c, const tsm_adcensus_config* cfg, const uint8_t* left, size_t lstep, const uint8_t* right,
                         size_t rstep, int H, int W)
{
    if (!c) return TSM_E_ARG;
    int rc = check_images(c, left, lstep, right, rstep, H, W);
    if (rc) return rc;
    if (c->pending) return fail(c, TSM_E_STATE, "tsm_adcensus_enqueue: previous pair not waited for");
    if ((rc = ensure_arena(c, cfg, H, W))) return rc;
    if ((rc = upload_pair_host(c, left, lstep, right, rstep, H, W))) return rc;
    c->have_pair = true;
    if ((rc = run_stages(c, TSM_STAGE_ALL, -1))) return rc;
    if ((rc = ensure_pinned(c, (void**)&c->h_out, &c->h_out_bytes, (size_t)H * W * 4))) return rc;
    CK(c, cudaMemcpyAsync(c->h_out, c->fin.p, (size_t)H * W * 4, cudaMemcpyDeviceToHost, c->stream));
    c->pending = true;
    return TSM_OK;
}

int tsm_adcensus_wait(tsm_ctx* c, float* disparity, size_t dstep)
{
    if (!c) return TSM_E_ARG;
    if (!c->pending) return fail(c, TSM_E_STATE, "tsm_adcensus_wait: nothing enqueued");
    const int H = c->dm.H, W = c->dm.W;
    if (!disparity || dstep < (size_t)W * 4) return fail(c, TSM_E_ARG, "[ADCensus] disparity buffer error.");
    c->pending = false;
    CK(c, cudaStreamSynchronize(c->stream));
    for (int y = 0; y < H; ++y) memcpy((uint8_t*)disparity + (size_t)y * dstep, c->h_out + (size_t)y * W, (size_t)W * 4);
    return TSM_OK;
}

int tsm_adcensus_compute(tsm_ctx* c, const tsm_adcensus_config* cfg, const uint8_t* left, size_t lstep, const uint8_t* right,
                         size_t rstep, int H, int W, float* disparity, size_t dstep)
{
    if (!c) return TSM_E_ARG;
    if (!disparity || dstep < (size_t)W * 4) return fail(c, TSM_E_ARG, "[ADCensus] disparity buffer error.");
    int rc = tsm_adcensus_enqueue(c, cfg, left, lstep, right, rstep, H, W);
    if (rc) return rc;
    return tsm_adcensus_wait(c, disparity, dstep);
}

// ------------------------------------------------------------------ rectify
static int upload_maps(tsm_ctx* c, int slot, const void* m1, const void* m2, int kind, int H, int W)
{
    const size_t n = (size_t)H * W;
    int rc;
    if ((rc = ensure(c, c->r_map1[slot], n * 4))) return rc;
    if ((rc = ensure(c, c->r_map2[slot], n * 2))) return rc;
    if (kind == TSM_MAP_FIXED_16SC2_16UC1) {
        CK(c, cudaMemcpyAsync(c->r_map1[slot].p, m1, n * 4, cudaMemcpyHostToDevice, c->stream));
        CK(c, cudaMemcpyAsync(c->r_map2[slot].p, m2, n * 2, cudaMemcpyHostToDevice, c->stream));
    } else if (kind == TSM_MAP_FLOAT_32FC1_X2) {
        if ((rc = ensure(c, c->r_fmap[slot][0], n * 4))) return rc;
        if ((rc = ensure(c, c->r_fmap[slot][1], n * 4))) return rc;
        CK(c, cudaMemcpyAsync(c->r_fmap[slot][0].p, m1, n * 4, cudaMemcpyHostToDevice, c->stream));
        CK(c, cudaMemcpyAsync(c->r_fmap[slot][1].p, m2, n * 4, cudaMemcpyHostToDevice, c->stream));
        Launcher L{c->stream, &c->launches};
        convert_maps(L, (const float*)c->r_fmap[slot][0].p, (const float*)c->r_fmap[slot][1].p, H, W,
                     (int16_t*)c->r_map1[slot].p, (uint16_t*)c->r_map2[slot].p);
    } else {
        return fail(c, TSM_E_ARG, "[EpipolarRectify] unknown map kind %d", kind);
    }
    // pageable host memory: the copies above are staged synchronously by the runtime.
    return TSM_OK;
}

// Device copy of one map pair in `slot`.  The copy is reused only while the caller vouches, through a non-zero
// generation id, that the CONTENT behind (m1, m2) is unchanged: host addresses alone say nothing (a reloaded map can
// land on the same address).  generation 0 = upload every call.
static int ensure_map_slot(tsm_ctx* c, int slot, unsigned long long gen, const void* m1, const void* m2, int kind, int H, int W)
{
    tsm_ctx::MapSlot& s = c->map_slot[slot];
    if (gen != 0 && s.gen == gen && s.m1 == m1 && s.m2 == m2 && s.kind == kind && s.H == H && s.W == W) return TSM_OK;
    s.gen = 0;
    int rc = upload_maps(c, slot, m1, m2, kind, H, W);
    if (rc) return rc;
    s.gen = gen; s.m1 = m1; s.m2 = m2; s.kind = kind; s.H = H; s.W = W;
    return TSM_OK;
}

int tsm_remap(tsm_ctx* c, const uint8_t* src, size_t sstep, int sH, int sW, const void* map1, const void* map2, int map_kind,
              unsigned long long map_generation, int H, int W, uint8_t* dst, size_t dstep)
{
    if (!c) return TSM_E_ARG;
    if (!map1 || !map2) return fail(c, TSM_E_ARG, "Stereo epipolar rectify params is empty, please load it first.");
    if (!src || sH <= 0 || sW <= 0 || sstep < (size_t)sW * 3) return fail(c, TSM_E_ARG, "Left or Right image is empty.");
    if (!dst || H <= 0 || W <= 0 || dstep < (size_t)W * 3) return fail(c, TSM_E_ARG, "[EpipolarRectify] destination error.");
    CK(c, cudaSetDevice(c->device));
    int rc;
    // a slot that already holds this pair, else the other slot than the one used last
    int slot = 0;
    for (int k = 0; k < 2; ++k)
        if (map_generation != 0 && c->map_slot[k].gen == map_generation && c->map_slot[k].m1 == map1 && c->map_slot[k].m2 == map2) slot = k;
    if (!(map_generation != 0 && c->map_slot[slot].gen == map_generation && c->map_slot[slot].m1 == map1)) {
        slot = c->map_slot[0].gen == 0 ? 0 : (c->map_slot[1].gen == 0 ? 1 : (c->remap_victim ^= 1));
    }
    if ((rc = ensure_map_slot(c, slot, map_generation, map1, map2, map_kind, H, W))) return rc;
    if ((rc = ensure(c, c->r_src, sstep * sH + (size_t)W * 3 * H))) return rc;
    uint8_t* d_src = (uint8_t*)c->r_src.p;
    uint8_t* d_dst = d_src + sstep * sH;
    CK(c, cudaMemcpyAsync(d_src, src, sstep * sH, cudaMemcpyHostToDevice, c->stream));
    Launcher L{c->stream, &c->launches};
    remap_bilinear(L, d_src, sstep, sH, sW, (const int16_t*)c->r_map1[slot].p, (const uint16_t*)c->r_map2[slot].p, H, W, d_dst,
                   (size_t)W * 3);
    CK(c, cudaGetLastError());
    CK(c, cudaMemcpy2DAsync(dst, dstep, d_dst, (size_t)W * 3, (size_t)W * 3, H, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    return TSM_OK;
}

void tsm_invalidate_maps(tsm_ctx* c)
{
    if (!c) return;
    for (auto& s : c->map_slot) s.gen = 0;
}

static int ensure_map_cache(tsm_ctx* c, const void* m00, const void* m01, const void* m10, const void* m11, int kind,
                            unsigned long long gen, int H, int W)
{
    if (!m00 || !m01 || !m10 || !m11)
        return fail(c, TSM_E_ARG, "Stereo epipolar rectify params is empty, please load it first.");
    int rc;
    if ((rc = ensure_map_slot(c, 0, gen, m00, m01, kind, H, W))) return rc;
    if ((rc = ensure_map_slot(c, 1, gen, m10, m11, kind, H, W))) return rc;
    return TSM_OK;
}

// Remaps both halves of a device-resident side-by-side frame into d_left / d_right (packed BGR).
static int rectify_on_device(tsm_ctx* c, const uint8_t* d_src, size_t row, int H, int W, const void* m00, const void* m01,
                             const void* m10, const void* m11, int kind, unsigned long long gen, uint8_t* d_left, uint8_t* d_right)
{
    int rc;
    if ((rc = ensure_map_cache(c, m00, m01, m10, m11, kind, gen, H, W))) return rc;
    Launcher L{c->stream, &c->launches};
    ScopedStage s(c, "rectify");
    // halves [0,W) and [W,2W) of the frame (EpipolarRectify.cpp:81-82): same rows, column offset 3*W bytes
    remap_bilinear(L, d_src, row, H, W, (const int16_t*)c->r_map1[0].p, (const uint16_t*)c->r_map2[0].p, H, W, d_left, (size_t)W * 3);
    remap_bilinear(L, d_src + (size_t)W * 3, row, H, W, (const int16_t*)c->r_map1[1].p, (const uint16_t*)c->r_map2[1].p, H, W,
                   d_right, (size_t)W * 3);
    CK(c, cudaGetLastError());
    return TSM_OK;
}

// Uploads the side-by-side frame and remaps both halves into d_left / d_right (packed BGR).
static int rectify_to_device(tsm_ctx* c, const uint8_t* stereo, size_t sstep, int H, int W, const void* m00, const void* m01,
                             const void* m10, const void* m11, int kind, unsigned long long gen, uint8_t* d_left, uint8_t* d_right)
{
    if (!stereo || H <= 0 || W <= 0 || sstep < (size_t)W * 6) return fail(c, TSM_E_ARG, "Stereo image is empty.");
    int rc;
    const size_t row = (size_t)W * 6;
    if ((rc = ensure_pinned(c, (void**)&c->h_stereo, &c->h_stereo_bytes, row * H))) return rc;
    for (int y = 0; y < H; ++y) memcpy(c->h_stereo + (size_t)y * row, stereo + (size_t)y * sstep, row);
    if ((rc = ensure(c, c->r_src, row * H + (size_t)W * 3 * H))) return rc;
    uint8_t* d_src = (uint8_t*)c->r_src.p;
    CK(c, cudaMemcpyAsync(d_src, c->h_stereo, row * H, cudaMemcpyHostToDevice, c->stream));
    return rectify_on_device(c, d_src, row, H, W, m00, m01, m10, m11, kind, gen, d_left, d_right);
}

int tsm_rectify_stereo(tsm_ctx* c, const uint8_t* stereo, size_t sstep, int H, int W, const void* map00, const void* map01,
                       const void* map10, const void* map11, int map_kind, unsigned long long map_generation, uint8_t* left,
                       size_t lstep, uint8_t* right, size_t rstep)
{
    if (!c) return TSM_E_ARG;
    if (!left || !right || lstep < (size_t)W * 3 || rstep < (size_t)W * 3)
        return fail(c, TSM_E_ARG, "[EpipolarRectify] destination error.");
    if (H <= 0 || W <= 0) return fail(c, TSM_E_ARG, "Stereo image is empty.");
    if (c->pending) return fail(c, TSM_E_STATE, "tsm_rectify_stereo: a pair is enqueued on this context and not waited for");
    CK(c, cudaSetDevice(c->device));  // before anything that may allocate
    int rc;
    const size_t img_bytes = (size_t)H * W * 3;
    if ((rc = ensure(c, c->img[0], img_bytes))) return rc;
    if ((rc = ensure(c, c->img[1], img_bytes))) return rc;
    if ((rc = rectify_to_device(c, stereo, sstep, H, W, map00, map01, map10, map11, map_kind, map_generation,
                                (uint8_t*)c->img[0].p, (uint8_t*)c->img[1].p)))
        return rc;
    CK(c, cudaMemcpy2DAsync(left, lstep, c->img[0].p, (size_t)W * 3, (size_t)W * 3, H, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaMemcpy2DAsync(right, rstep, c->img[1].p, (size_t)W * 3, (size_t)W * 3, H, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    return TSM_OK;
}

int tsm_rectify_adcensus_enqueue(tsm_ctx* c, const tsm_adcensus_config* cfg, const uint8_t* stereo, size_t sstep, int H, int W,
                                 const void* map00, const void* map01, const void* map10, const void* map11, int map_kind,
                                 unsigned long long map_generation)
{
    if (!c) return TSM_E_ARG;
    if (c->pending) return fail(c, TSM_E_STATE, "tsm_rectify_adcensus: previous pair not waited for");
    int rc = ensure_arena(c, cfg, H, W);
    if (rc) return rc;
    if (c->profiling) c->timers_used = 0;
    if ((rc = rectify_to_device(c, stereo, sstep, H, W, map00, map01, map10, map11, map_kind, map_generation,
                                (uint8_t*)c->img[0].p, (uint8_t*)c->img[1].p)))
        return rc;
    c->have_pair = true;
    if ((rc = run_stages(c, TSM_STAGE_ALL, -2))) return rc;
    if ((rc = ensure_pinned(c, (void**)&c->h_out, &c->h_out_bytes, (size_t)H * W * 4))) return rc;
    CK(c, cudaMemcpyAsync(c->h_out, c->fin.p, (size_t)H * W * 4, cudaMemcpyDeviceToHost, c->stream));
    c->pending = true;
    return TSM_OK;
}

int tsm_rectify_adcensus(tsm_ctx* c, const tsm_adcensus_config* cfg, const uint8_t* stereo, size_t sstep, int H, int W,
                         const void* map00, const void* map01, const void* map10, const void* map11, int map_kind,
                         unsigned long long map_generation, float* disparity, size_t dstep)
{
    if (!c) return TSM_E_ARG;
    if (!disparity || dstep < (size_t)W * 4) return fail(c, TSM_E_ARG, "[ADCensus] disparity buffer error.");
    int rc = tsm_rectify_adcensus_enqueue(c, cfg, stereo, sstep, H, W, map00, map01, map10, map11, map_kind, map_generation);
    if (rc) return rc;
    return tsm_adcensus_wait(c, disparity, dstep);
}

int tsm_rectify_adcensus_device(tsm_ctx* c, const tsm_adcensus_config* cfg, const uint8_t* d_stereo, size_t sstep, int H, int W,
                                const void* map00, const void* map01, const void* map10, const void* map11, int map_kind,
                                unsigned long long map_generation, float* d_disparity)
{
    if (!c) return TSM_E_ARG;
    if (!d_stereo || !d_disparity || H <= 0 || W <= 0 || sstep < (size_t)W * 6) return fail(c, TSM_E_ARG, "Stereo image is empty.");
    if (c->pending) return fail(c, TSM_E_STATE, "tsm_rectify_adcensus_device: previous pair not waited for");
    int rc = ensure_arena(c, cfg, H, W);
    if (rc) return rc;
    if (c->profiling) c->timers_used = 0;
    if ((rc = rectify_on_device(c, d_stereo, sstep, H, W, map00, map01, map10, map11, map_kind, map_generation,
                                (uint8_t*)c->img[0].p, (uint8_t*)c->img[1].p)))
        return rc;
    c->have_pair = true;
    if ((rc = run_stages(c, TSM_STAGE_ALL, -2))) return rc;
    CK(c, cudaMemcpyAsync(d_disparity, c->fin.p, (size_t)H * W * 4, cudaMemcpyDeviceToDevice, c->stream));
    return TSM_OK;
}

// --------------------------------------------------------------------- taps
int tsm_stage_begin(tsm_ctx* c, const tsm_adcensus_config* cfg, const uint8_t* left, size_t lstep, const uint8_t* right,
                    size_t rstep, int H, int W)
{
    if (!c) return TSM_E_ARG;
    int rc = check_images(c, left, lstep, right, rstep, H, W);
    if (rc) return rc;
    if ((rc = ensure_arena(c, cfg, H, W))) return rc;
    if ((rc = upload_pair_host(c, left, lstep, right, rstep, H, W))) return rc;
    CK(c, cudaStreamSynchronize(c->stream));
    c->have_pair = true;
    c->disp_cur = 0;
    return TSM_OK;
}

int tsm_stage_run(tsm_ctx* c, int mask, int arg)
{
    if (!c) return TSM_E_ARG;
    if (!c->have_pair) return fail(c, TSM_E_STATE, "tsm_stage_run: call tsm_stage_begin first");
    CK(c, cudaSetDevice(c->device));
    c->stage_mode = true;
    int rc = run_stages(c, mask, arg);
    c->stage_mode = false;
    if (rc) return rc;
    CK(c, cudaStreamSynchronize(c->stream));
    return TSM_OK;
}

int tsm_volume_pitch(const tsm_ctx* c) { return c ? c->dm.Dn : 0; }

static Buf* tap_buffer(tsm_ctx* c, int id, size_t* bytes)
{
    const size_t npx = c->dm.npx();
    switch (id) {
        case TSM_BUF_VOL_LEFT: *bytes = npx * c->dm.Dn * 4; return &c->vol[0];
        case TSM_BUF_VOL_RIGHT: *bytes = npx * c->dm.Dn * 4; return &c->vol[1];
        case TSM_BUF_ARMS_LEFT: *bytes = npx * 4; return &c->arms[0];
        case TSM_BUF_ARMS_RIGHT: *bytes = npx * 4; return &c->arms[1];
        case TSM_BUF_WTA_LEFT: *bytes = npx * 4; return &c->wta_[0];
        case TSM_BUF_WTA_RIGHT: *bytes = npx * 4; return &c->wta_[1];
        case TSM_BUF_DISP: *bytes = npx * 4; return &c->disp[c->disp_cur];
        case TSM_BUF_EDGES: *bytes = npx; return &c->e_edges;
        case TSM_BUF_FINAL: *bytes = npx * 4; return &c->fin;
        case TSM_BUF_CENSUS_LEFT: *bytes = npx * 48; return &c->census[0];
        case TSM_BUF_CENSUS_RIGHT: *bytes = npx * 48; return &c->census[1];
        case TSM_BUF_IMG_LEFT: *bytes = npx * 3; return &c->img[0];
        case TSM_BUF_IMG_RIGHT: *bytes = npx * 3; return &c->img[1];
        case TSM_BUF_IMG4_LEFT: *bytes = npx * 4; return &c->img4[0];
        case TSM_BUF_IMG4_RIGHT: *bytes = npx * 4; return &c->img4[1];
        default: return nullptr;
    }
}

size_t tsm_buffer_bytes(const tsm_ctx* c, int buffer)
{
    if (!c) return 0;
    size_t bytes = 0;
    return tap_buffer(const_cast<tsm_ctx*>(c), buffer, &bytes) ? bytes : 0;
}

int tsm_tap(tsm_ctx* c, int buffer, void* dst, size_t bytes)
{
    if (!c || !dst) return TSM_E_ARG;
    if (!c->have_pair) return fail(c, TSM_E_STATE, "tsm_tap: nothing computed yet");
    size_t need = 0;
    Buf* b = tap_buffer(c, buffer, &need);
    if (!b || !b->p) return fail(c, TSM_E_ARG, "tsm_tap: unknown buffer %d", buffer);
    if (bytes != need) return fail(c, TSM_E_ARG, "tsm_tap: buffer %d holds %zu bytes, caller passed %zu", buffer, need, bytes);
    CK(c, cudaSetDevice(c->device));
    const void* srcp = b->p;
    if (buffer == TSM_BUF_VOL_LEFT || buffer == TSM_BUF_VOL_RIGHT) {  // split layout -> dense [H][W][Dn]
        int rc = ensure(c, c->dense, need);
        if (rc) return rc;
        Launcher L{c->stream, &c->launches};
        volume_gather(L, c->dm, view_ptrs(c, buffer - TSM_BUF_VOL_LEFT).vol, (float*)c->dense.p);
        srcp = c->dense.p;
    }
    CK(c, cudaMemcpyAsync(dst, srcp, need, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    return TSM_OK;
}

int tsm_poke(tsm_ctx* c, int buffer, const void* src, size_t bytes)
{
    if (!c || !src) return TSM_E_ARG;
    if (!c->have_pair) return fail(c, TSM_E_STATE, "tsm_poke: call tsm_stage_begin first");
    size_t need = 0;
    Buf* b = tap_buffer(c, buffer, &need);
    if (!b || !b->p) return fail(c, TSM_E_ARG, "tsm_poke: unknown buffer %d", buffer);
    if (bytes != need) return fail(c, TSM_E_ARG, "tsm_poke: buffer %d holds %zu bytes, caller passed %zu", buffer, need, bytes);
    CK(c, cudaSetDevice(c->device));
    if (buffer == TSM_BUF_VOL_LEFT || buffer == TSM_BUF_VOL_RIGHT) {  // dense [H][W][Dn] -> split layout
        int rc = ensure(c, c->dense, need);
        if (rc) return rc;
        CK(c, cudaMemcpyAsync(c->dense.p, src, need, cudaMemcpyHostToDevice, c->stream));
        Launcher L{c->stream, &c->launches};
        volume_scatter(L, c->dm, (const float*)c->dense.p, view_ptrs(c, buffer - TSM_BUF_VOL_LEFT).vol);
        CK(c, cudaStreamSynchronize(c->stream));
        return TSM_OK;
    }
    CK(c, cudaMemcpyAsync(b->p, src, need, cudaMemcpyHostToDevice, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    return TSM_OK;
}

// ---- rectify-map generation ----------------------------------------------------------------------
static bool invert3x3(const double* m, double* inv)
{
    // LU with partial pivoting on [m | I] (cv::invert DECOMP_LU)
    double a[3][6];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) { a[r][c] = m[3 * r + c]; a[r][3 + c] = r == c ? 1.0 : 0.0; }
    for (int col = 0; col < 3; ++col) {
        int piv = col;
        for (int r = col + 1; r < 3; ++r)
            if (fabs(a[r][col]) > fabs(a[piv][col])) piv = r;
        if (fabs(a[piv][col]) < 1e-300) return false;
        if (piv != col)
            for (int c = 0; c < 6; ++c) std::swap(a[piv][c], a[col][c]);
        const double d = 1.0 / a[col][col];
        for (int r = col + 1; r < 3; ++r) {
            const double f = a[r][col] * d;
            for (int c = col; c < 6; ++c) a[r][c] -= f * a[col][c];
        }
    }
    for (int c = 3; c < 6; ++c)
        for (int r = 2; r >= 0; --r) {
            double s = a[r][c];
            for (int k = r + 1; k < 3; ++k) s -= a[r][k] * a[k][c];
            a[r][c] = s / a[r][r];
        }
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) inv[3 * r + c] = a[r][3 + c];
    return true;
}

int tsm_init_undistort_rectify_map(tsm_ctx* c, const double* K, const double* dist, int n_dist, const double* R,
                                   const double* newK, int ld_new, int H, int W, int16_t* map1, size_t step1, uint16_t* map2,
                                   size_t step2)
{
    if (!c) return TSM_E_ARG;
    if (!K || H <= 0 || W <= 0 || !map1 || !map2 || step1 < (size_t)W * 4 || step2 < (size_t)W * 2)
        return fail(c, TSM_E_ARG, "tsm_init_undistort_rectify_map: bad argument");
    if (!(n_dist == 0 || n_dist == 4 || n_dist == 5 || n_dist == 8 || n_dist == 12 || n_dist == 14) || (n_dist && !dist))
        return fail(c, TSM_E_ARG, "tsm_init_undistort_rectify_map: distortion coefficients must be 0, 4, 5, 8, 12 or 14 values");
    if (newK && ld_new != 3 && ld_new != 4) return fail(c, TSM_E_ARG, "tsm_init_undistort_rectify_map: ld_new must be 3 or 4");
    CK(c, cudaSetDevice(c->device));
    double Ar[9], Rm[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, ArR[9], ir[9], k[14] = {0}, tilt[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    for (int r = 0; r < 3; ++r)
        for (int q = 0; q < 3; ++q) Ar[3 * r + q] = newK ? newK[ld_new * r + q] : K[3 * r + q];
    if (!newK) {  // cv::getDefaultNewCameraMatrix(K, size, centerPrincipalPoint = true)
        Ar[2] = (W - 1) * 0.5;
        Ar[5] = (H - 1) * 0.5;
    }
    if (R) memcpy(Rm, R, sizeof Rm);
    for (int r = 0; r < 3; ++r)
        for (int q = 0; q < 3; ++q) {
            double s = 0;
            for (int t = 0; t < 3; ++t) s += Ar[3 * r + t] * Rm[3 * t + q];
            ArR[3 * r + q] = s;
        }
    if (!invert3x3(ArR, ir)) return fail(c, TSM_E_ARG, "tsm_init_undistort_rectify_map: new camera matrix * R is singular");
    for (int i = 0; i < n_dist; ++i) k[i] = dist[i];
    if (k[12] != 0 || k[13] != 0) {
        // computeTiltProjectionMatrix(tauX, tauY): matTilt = projZ(Ry Rx) * Ry Rx
        const double cx_ = cos(k[12]), sx_ = sin(k[12]), cy_ = cos(k[13]), sy_ = sin(k[13]);
        const double rx[9] = {1, 0, 0, 0, cx_, sx_, 0, -sx_, cx_}, ry[9] = {cy_, 0, -sy_, 0, 1, 0, sy_, 0, cy_};
        double rxy[9];
        for (int r = 0; r < 3; ++r)
            for (int q = 0; q < 3; ++q) {
                double s = 0;
                for (int t = 0; t < 3; ++t) s += ry[3 * r + t] * rx[3 * t + q];
                rxy[3 * r + q] = s;
            }
        const double pz[9] = {rxy[8], 0, -rxy[2], 0, rxy[8], -rxy[5], 0, 0, 1};
        for (int r = 0; r < 3; ++r)
            for (int q = 0; q < 3; ++q) {
                double s = 0;
                for (int t = 0; t < 3; ++t) s += pz[3 * r + t] * rxy[3 * t + q];
                tilt[3 * r + q] = s;
            }
    }
    int rc;
    const size_t n = (size_t)H * W;
    if ((rc = ensure(c, c->k_out, n * 12))) return rc;
    int16_t* d1 = (int16_t*)c->k_out.p;
    uint16_t* d2 = (uint16_t*)(d1 + 2 * n);
    Launcher L{c->stream, &c->launches};
    init_undistort_rectify_map(L, ir, k, tilt, K[0], K[4], K[2], K[5], H, W, d1, d2);
    CK(c, cudaGetLastError());
    CK(c, cudaMemcpy2DAsync(map1, step1, d1, (size_t)W * 4, (size_t)W * 4, H, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaMemcpy2DAsync(map2, step2, d2, (size_t)W * 2, (size_t)W * 2, H, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    return TSM_OK;
}

// ---- disparity consumers -------------------------------------------------------------------------
// Resolves the input map: a host map is staged densely into k_in; NULL selects the context's last result.
static int consumer_input(tsm_ctx* c, const float* disparity, size_t step, int H, int W, const float** d_in)
{
    if (H <= 0 || W <= 0) return fail(c, TSM_E_ARG, "disparity consumer: empty map");
    const size_t n = (size_t)H * W;
    if (!disparity) {
        if (c->fin_H != H || c->fin_W != W || !c->fin.p)
            return fail(c, TSM_E_STATE, "disparity consumer: no %dx%d disparity map of this context to consume", W, H);
        *d_in = (const float*)c->fin.p;
        return TSM_OK;
    }
    if (step < (size_t)W * 4) return fail(c, TSM_E_ARG, "disparity consumer: row stride too small");
    int rc;
    if ((rc = ensure(c, c->k_in, n * 4))) return rc;
    CK(c, cudaMemcpy2DAsync(c->k_in.p, (size_t)W * 4, disparity, step, (size_t)W * 4, H, cudaMemcpyHostToDevice, c->stream));
    *d_in = (const float*)c->k_in.p;
    return TSM_OK;
}

int tsm_reproject_to_depth(tsm_ctx* c, const float* disparity, size_t step, int H, int W, float focal, float baseline, float* depth,
                           size_t dstep)
{
    if (!c) return TSM_E_ARG;
    if (!depth || dstep < (size_t)W * 4) return fail(c, TSM_E_ARG, "tsm_reproject_to_depth: destination error");
    CK(c, cudaSetDevice(c->device));
    const float* d_in = nullptr;
    int rc;
    if ((rc = consumer_input(c, disparity, step, H, W, &d_in))) return rc;
    const size_t n = (size_t)H * W;
    if ((rc = ensure(c, c->k_out, n * 12))) return rc;
    Launcher L{c->stream, &c->launches};
    reproject_depth(L, d_in, (float*)c->k_out.p, n, focal, baseline);
    CK(c, cudaGetLastError());
    CK(c, cudaMemcpy2DAsync(depth, dstep, c->k_out.p, (size_t)W * 4, (size_t)W * 4, H, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    return TSM_OK;
}

static int reproject_3d_common(tsm_ctx* c, const float* disparity, size_t step, int H, int W, float* xyz, size_t xstep,
                               const double* Q, float focal, float baseline, float cx, float cy)
{
    if (!xyz || xstep < (size_t)W * 12) return fail(c, TSM_E_ARG, "tsm_reproject_to_3d: destination error");
    CK(c, cudaSetDevice(c->device));
    const float* d_in = nullptr;
    int rc;
    if ((rc = consumer_input(c, disparity, step, H, W, &d_in))) return rc;
    const size_t n = (size_t)H * W;
    if ((rc = ensure(c, c->k_out, n * 12))) return rc;
    Launcher L{c->stream, &c->launches};
    if (Q) reproject_xyz_q(L, d_in, (float*)c->k_out.p, H, W, Q);
    else reproject_xyz_fb(L, d_in, (float*)c->k_out.p, H, W, focal, baseline, cx, cy);
    CK(c, cudaGetLastError());
    CK(c, cudaMemcpy2DAsync(xyz, xstep, c->k_out.p, (size_t)W * 12, (size_t)W * 12, H, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    return TSM_OK;
}

int tsm_reproject_to_3d(tsm_ctx* c, const float* disparity, size_t step, int H, int W, float focal, float baseline, float cx, float cy,
                        float* xyz, size_t xstep)
{
    if (!c) return TSM_E_ARG;
    return reproject_3d_common(c, disparity, step, H, W, xyz, xstep, nullptr, focal, baseline, cx, cy);
}

int tsm_reproject_to_3d_q(tsm_ctx* c, const float* disparity, size_t step, int H, int W, const double* Q, float* xyz, size_t xstep)
{
    if (!c) return TSM_E_ARG;
    if (!Q) return fail(c, TSM_E_ARG, "tsm_reproject_to_3d_q: Q is empty");
    return reproject_3d_common(c, disparity, step, H, W, xyz, xstep, Q, 0.f, 0.f, 0.f, 0.f);
}

int tsm_apply_colormap(tsm_ctx* c, const float* disparity, size_t step, int H, int W, int auto_range, float min_val, float max_val,
                       const uint8_t* colormap, uint8_t* dst, size_t dstep)
{
    if (!c) return TSM_E_ARG;
    if (!dst || dstep < (size_t)W * 3) return fail(c, TSM_E_ARG, "tsm_apply_colormap: destination error");
    CK(c, cudaSetDevice(c->device));
    const float* d_in = nullptr;
    int rc;
    if ((rc = consumer_input(c, disparity, step, H, W, &d_in))) return rc;
    const size_t n = (size_t)H * W;
    if ((rc = ensure(c, c->k_out, n * 12))) return rc;
    if ((rc = ensure(c, c->k_tab, 768))) return rc;
    if ((rc = ensure(c, c->k_range, 8))) return rc;
    uint8_t jet[768];
    if (!colormap) {
        jet_colormap(jet);
        colormap = jet;
    }
    CK(c, cudaMemcpyAsync(c->k_tab.p, colormap, 768, cudaMemcpyHostToDevice, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));  // `jet` lives on this stack frame
    Launcher L{c->stream, &c->launches};
    apply_colormap(L, d_in, (uint8_t*)c->k_out.p, n, auto_range != 0, min_val, max_val, (const uint8_t*)c->k_tab.p,
                   (int*)c->k_range.p);
    CK(c, cudaGetLastError());
    CK(c, cudaMemcpy2DAsync(dst, dstep, c->k_out.p, (size_t)W * 3, (size_t)W * 3, H, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    return TSM_OK;
}

void tsm_jet_colormap(uint8_t* table768)
{
    if (table768) jet_colormap(table768);
}

// writePointCloudToPCD / writePointCloudToPLY (stereo.cpp:204-356): same headers, same number formatting (std::to_chars),
// same point filter.  Host-only.
int tsm_write_point_cloud(const uint8_t* bgr, size_t cstep, const float* xyz, size_t xstep, int H, int W, const char* path,
                          int format, size_t* points)
{
    if (!bgr || !xyz || !path || !path[0] || H <= 0 || W <= 0 || cstep < (size_t)W * 3 || xstep < (size_t)W * 12)
        return fail(nullptr, TSM_E_ARG, "Empty input.");  // the reference logs "Empty input." and returns
    if (format != TSM_CLOUD_PCD && format != TSM_CLOUD_PLY) return fail(nullptr, TSM_E_ARG, "tsm_write_point_cloud: unknown format %d", format);
    const float inf = std::numeric_limits<float>::infinity();
    size_t n = 0;
    for (int y = 0; y < H; ++y) {
        const float* row = (const float*)((const uint8_t*)xyz + (size_t)y * xstep);
        for (int x = 0; x < W; ++x)
            if (!(row[3 * x] == inf || row[3 * x + 1] == inf || row[3 * x + 2] == inf)) ++n;
    }
    std::string header;
    if (format == TSM_CLOUD_PCD) {
        header += "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z rgb\nSIZE 4 4 4 4\nTYPE F F F U\nCOUNT 1 1 1 1\n";
        header += "WIDTH " + std::to_string(n) + "\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS " + std::to_string(n) + "\nDATA ascii\n";
    } else {
        header += "ply\nformat ascii 1.0\nelement vertex " + std::to_string(n) + "\n";
        header += "property float x\nproperty float y\nproperty float z\nproperty uchar red\nproperty uchar green\nproperty uchar blue\nend_header\n";
    }
    const size_t cap = header.size() + 128 * n + 16;
    std::unique_ptr<char[]> buf(new (std::nothrow) char[cap]);
    if (!buf) return fail(nullptr, TSM_E_OOM, "tsm_write_point_cloud: out of host memory");
    char* cur = buf.get();
    char* const end = buf.get() + cap;
    memcpy(cur, header.data(), header.size());
    cur += header.size();
    for (int y = 0; y < H; ++y) {
        const float* prow = (const float*)((const uint8_t*)xyz + (size_t)y * xstep);
        const uint8_t* crow = bgr + (size_t)y * cstep;
        for (int x = 0; x < W; ++x) {
            const float* p = prow + 3 * x;
            if (p[0] == inf || p[1] == inf || p[2] == inf) continue;
            const uint8_t* c = crow + 3 * x;
            for (int k = 0; k < 3; ++k) {
                cur = std::to_chars(cur, end, p[k]).ptr;
                *cur++ = ' ';
            }
            if (format == TSM_CLOUD_PCD) {
                const unsigned rgb = (unsigned)c[2] << 16 | (unsigned)c[1] << 8 | (unsigned)c[0] | 1u << 24;
                cur = std::to_chars(cur, end, rgb).ptr;
            } else {
                cur = std::to_chars(cur, end, (int)c[2]).ptr;
                *cur++ = ' ';
                cur = std::to_chars(cur, end, (int)c[1]).ptr;
                *cur++ = ' ';
                cur = std::to_chars(cur, end, (int)c[0]).ptr;
            }
            *cur++ = '\n';
        }
    }
    FILE* f = fopen(path, "wb");  // binary mode: no newline translation (:248)
    if (!f) return fail(nullptr, TSM_E_ARG, "tsm_write_point_cloud: cannot open %s", path);
    const size_t want = (size_t)(cur - buf.get());
    const size_t wrote = fwrite(buf.get(), 1, want, f);
    fclose(f);
    if (wrote != want) return fail(nullptr, TSM_E_ARG, "tsm_write_point_cloud: short write to %s", path);
    if (points) *points = n;
    return TSM_OK;
}

int tsm_set_profiling(tsm_ctx* c, int enabled)
{
    if (!c) return TSM_E_ARG;
    c->profiling = enabled != 0;
    c->timers_used = 0;
    return TSM_OK;
}

int tsm_get_stage_times(tsm_ctx* c, int* n, const char** names, float* ms)
{
    if (!c || !n) return TSM_E_ARG;
    CK(c, cudaStreamSynchronize(c->stream));
    int filled = 0;
    for (size_t i = 0; i < c->timers_used && filled < *n; ++i) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, c->timers[i].beg, c->timers[i].end) != cudaSuccess) continue;
        if (names) names[filled] = c->timers[i].name;
        if (ms) ms[filled] = t;
        ++filled;
    }
    *n = filled;
    return TSM_OK;
}

long long tsm_launch_count(const tsm_ctx* c) { return c ? c->launches : 0; }

int tsm_selftest(tsm_ctx* c, int which, unsigned long long* mismatches)
{
    if (!c || !mismatches) return TSM_E_ARG;
    if (which != TSM_SELFTEST_DIV) return fail(c, TSM_E_ARG, "tsm_selftest: unknown test %d", which);
    CK(c, cudaSetDevice(c->device));
    unsigned long long* d = nullptr;
    CK(c, cudaMalloc(&d, sizeof *d));
    cudaMemsetAsync(d, 0, sizeof *d, c->stream);
    Launcher L{c->stream, &c->launches};
    selftest_div(L, d);
    cudaError_t e = cudaMemcpyAsync(mismatches, d, sizeof *d, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d);
    if (e != cudaSuccess) return fail(c, TSM_E_CUDA, "tsm_selftest: %s", cudaGetErrorString(e));
    return TSM_OK;
}

}  // extern "C"
