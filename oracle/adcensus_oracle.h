/*
 * oracle/adcensus_oracle.h -- TEST INFRASTRUCTURE (oracle), not product code.
 *
 * Plain-C CPU restatement of the reference's ADCensus path
 * (/root/reference/source/ADCensus.cpp, RGB colour model, roi=mask=false,
 * minDisparity = 0) used as the parity checker for the CUDA kernels and as the
 * "port" CPU baseline.  Parity is PINNED: tests/test_oracle_vs_ref.py checks
 * every stage of this restatement bit-for-bit against the unmodified reference
 * compiled into oracle/_ref/libadcensus_ref.so (oracle/ref_driver.cpp), and
 * tests/golden/ holds reference-generated vectors for boxes without
 * /root/reference.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may link or call this.
 *
 * Layout: volumes are [H][W][Dn] fp32 (d innermost) -- the reference keeps
 * [Dn][H][W] planes (ADCensus.cpp:289); values are identical, only the index
 * order differs.  Maps are [H][W].  Images are packed BGR [H][W][3] uint8.
 */
#ifndef ORACLE_ADCENSUS_ORACLE_H
#define ORACLE_ADCENSUS_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* RGB tunables, source/stereo_utils.cpp:271-326 */
#define ORC_LAMBDA_AD 10.f
#define ORC_LAMBDA_CENSUS 30.f
#define ORC_CENSUS_W 9
#define ORC_CENSUS_H 7
#define ORC_TAU1 20
#define ORC_TAU2 6
#define ORC_L1 34
#define ORC_L2 17
#define ORC_ITERATIONS 4
#define ORC_COLOR_DIFF 15
#define ORC_PI1 1.f
#define ORC_PI2 3.f
#define ORC_DISP_TOL 0
#define ORC_VOTING_THRESH 20
#define ORC_VOTING_RATIO 0.4f
#define ORC_MAX_SEARCH_DEPTH 20
#define ORC_CANNY_LOW 30
#define ORC_CANNY_HIGH 90
#define ORC_OCCLUSION (-1)
#define ORC_MISMATCH (-2)

/* a4: census sign planes. lt/gt: [H][W][3] u64, bit i (row-major 7x9 window)
 * = neighbour i < / > centre.  Border pixels (window not inside) get 0. */
void orc_census_signatures(const uint8_t* img, int H, int W, uint64_t* lt, uint64_t* gt);
/* a4 direct form (ADCensus.cpp:454-474) for one pixel pair; small cross-checks only. */
int orc_census_direct(const uint8_t* L, const uint8_t* R, int W, int y, int xl, int xr);
/* a3 integer AD sum |dB|+|dG|+|dR| (ADCensus.cpp:426-437 before the /3.f). */
int orc_ad3(const uint8_t* L, const uint8_t* R, int W, int y, int xl, int xr);
/* a5: both initial cost volumes (ADCensus.cpp:500-581). */
void orc_cost_init(const uint8_t* L, const uint8_t* R, int H, int W, int Dn, float* vol0, float* vol1);
/* a6: arm lengths of one image, 4 int32 maps up/down/left/right (ADCensus.cpp:604-683). */
void orc_arms(const uint8_t* img, int H, int W, int32_t* up, int32_t* down, int32_t* left, int32_t* right);
/* a7: 4-iteration cross-based aggregation of one volume, in place, exact fp32
 * summation order (ADCensus.cpp:685-793). */
void orc_aggregate(float* vol, int H, int W, int Dn, const int32_t* up, const int32_t* down,
                   const int32_t* left, const int32_t* right);
/* a8: 4 cascaded in-place scanline passes for one view (ADCensus.cpp:795-1011),
 * sequential semantics.  view 0: own = left image; view 1: own = right image. */
void orc_scanline(float* vol, int H, int W, int Dn, const uint8_t* own, const uint8_t* other, int view);
/* a9: WTA (ADCensus.cpp:1394-1413). */
void orc_wta(const float* vol, int H, int W, int Dn, int32_t* disp);
/* a10: left-right check (ADCensus.cpp:1013-1044). */
void orc_lrc(const int32_t* dl, const int32_t* dr, int H, int W, int maxD, int32_t* out);
/* a11: one regionVoting call incl. the histogram leak (ADCensus.cpp:1046-1159). */
void orc_region_voting(int32_t* disp, int H, int W, int Dn, const int32_t* up, const int32_t* down,
                       const int32_t* left, const int32_t* right, int horizontal_first);
/* a12 (ADCensus.cpp:1161-1239). */
void orc_proper_interpolation(int32_t* disp, int H, int W, const uint8_t* left_img);
/* a13 (ADCensus.cpp:1241-1342); vol = post-scanline LEFT volume. edges_out optional [H][W]. */
void orc_discontinuity_adjustment(int32_t* disp, int H, int W, int Dn, const float* vol, uint8_t* edges_out);
/* a14 (ADCensus.cpp:1344-1374). */
void orc_subpixel(const int32_t* disp, int H, int W, int Dn, const float* vol, float* out);

/* Optional taps of the full pipeline (NULL = skip). Same meaning as RefTaps in ref_driver.cpp
 * but volumes are [H][W][Dn]. */
typedef struct OrcTaps {
    float* vol_init[2];
    float* vol_agg[2];
    float* vol_scan[2];
    int32_t* arms[2][4];
    int32_t* wta[2];
    int32_t* lrc;
    int32_t* vote[5];
    int32_t* interp;
    int32_t* discont;
    float* final_disp;
    double t_init, t_agg, t_scan, t_multi;
} OrcTaps;

/* a1/a15: the whole path, ADCensus::compute (ADCensus.cpp:330-407) with RGB, D = 0..maxD. */
int orc_adcensus(const uint8_t* left, const uint8_t* right, int H, int W, int maxD, OrcTaps* taps);
int orc_omp_max_threads(void);
void orc_omp_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
