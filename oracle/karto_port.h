/* ORACLE / TEST INFRASTRUCTURE -- plain-C restatement of the reference's correlative scan
 * matcher and occupancy grid (karto::OccupancyGrid::CreateFromScans, Karto.h:5946-6274; karto::ScanMatcher, /root/reference/lib/karto_sdk/src/Mapper.cpp:477-1208 and the
 * Karto.h/Mapper.h types it uses).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this; the product library never does.
 *
 * Pinned against the reference itself: tests/test_oracle_vs_ref.py compares every function
 * here bit-for-bit with oracle/_ref/libkarto_ref.so (the unmodified reference compiled by
 * oracle/Makefile) and against the committed fixtures in tests/golden/ generated from it.
 */
#ifndef KARTO_PORT_H
#define KARTO_PORT_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct kp_params {
  double search_size;            /* ScanMatcher::Create searchSize (m)            Mapper.cpp:477 */
  double resolution;             /* grid resolution (m)                                           */
  double smear_deviation;        /* CorrelationGrid smear sigma (m)               Mapper.h:1213  */
  double range_threshold;        /* laser range threshold -> grid margin          Mapper.cpp:503 */
  double coarse_search_angle_offset;   /* Mapper::m_pCoarseSearchAngleOffset     Mapper.cpp:590 */
  double coarse_angle_resolution;      /* Mapper::m_pCoarseAngleResolution       Mapper.cpp:591 */
  double fine_search_angle_offset;     /* Mapper::m_pFineSearchAngleOffset       Mapper.cpp:627 */
  double distance_variance_penalty;    /* stored (already squared) value         Mapper.cpp:675 */
  double angle_variance_penalty;       /* stored (already squared) value         Mapper.cpp:681 */
  double minimum_distance_penalty;
  double minimum_angle_penalty;
  int32_t use_response_expansion;      /* Mapper.cpp:594 */
} kp_params;

typedef struct kp_scan {
  int32_t n;                 /* number of range readings                                   */
  const double * ranges;     /* n raw range readings (NaN / Inf allowed)                   */
  const double * points_xy;  /* n UNFILTERED point readings, world frame, x,y interleaved  */
  double sensor_pose[3];     /* LocalizedRangeScan::GetSensorPose()                        */
} kp_scan;

typedef struct kp_matcher kp_matcher;

/* LocalizedRangeScan::Update (Karto.h:5644-5704): unfiltered point readings from ranges */
void kp_point_readings(const double * ranges, int32_t n, const double sensor_pose[3],
                       double minimum_angle, double angular_resolution, double * points_xy);

/* ScanMatcher::Create (Mapper.cpp:477-522); NULL on invalid parameters */
kp_matcher * kp_create(const kp_params * p);
void kp_destroy(kp_matcher * m);

/* ScanMatcher::MatchScan (Mapper.cpp:534-639); returns response, fills mean[3], cov[9] (row major) */
double kp_match(kp_matcher * m, const kp_scan * query, const kp_scan * base, int32_t nbase,
                int32_t do_penalize, int32_t do_refine, double mean[3], double cov[9]);

/* raster only: centre grid on query (Mapper.cpp:560-569) + AddScans (Mapper.cpp:1032-1105) */
void kp_raster(kp_matcher * m, const kp_scan * query, const kp_scan * base, int32_t nbase);

/* ScanMatcher::CorrelateScan (Mapper.cpp:712-862) on the grid as last rasterised.
 * If sums != NULL it receives the integer response volume (nY*nX*nAngles, index (y*nX+x)*nA+a);
 * dims[3] = {nX, nY, nAngles}.  cov is in/out (fine pass keeps the coarse xy terms). */
double kp_correlate(kp_matcher * m, const kp_scan * query, const double center[3],
                    const double sp_off[2], const double sp_res[2], double ang_off, double ang_res,
                    int32_t do_penalize, int32_t fine, double mean[3], double cov[9],
                    int32_t * sums, int32_t sums_cap, int32_t dims[3]);

/* ScanMatcher::FindValidPoints (Mapper.cpp:1113-1164); returns count */
int32_t kp_find_valid_points(const kp_scan * scan, const double viewpoint[2], double * out_xy);

/* GridIndexLookup::ComputeOffsets (Karto.h:6797-6894) with the grid's current offset.
 * out = nAngles x n int32 (row stride n).  Returns nAngles. */
int32_t kp_offsets(kp_matcher * m, const kp_scan * query, double angle_center, double angle_offset,
                   double angle_res, int32_t * out);

/* info: width,height,stride,roi_x,roi_y,roi_w,roi_h,data_size,kernel_size */
const uint8_t * kp_grid(kp_matcher * m, int32_t info[9], double off[2]);
const uint8_t * kp_kernel(kp_matcher * m);

/* karto::OccupancyGrid::CreateFromScans (Karto.h:5946-5961): hit / pass counters by Bresenham traces
 * (Karto.h:4874-4927, 6139-6229), then cell states 0 unknown / 100 occupied / 255 free (Karto.h:6241-6274).
 * NULL for n <= 0.  Laser limits are those of the scans' LaserRangeFinder. */
typedef struct kp_occupancy kp_occupancy;
kp_occupancy * kp_occupancy_create(const kp_scan * scans, int32_t n, double resolution, double range_threshold,
                                   double minimum_range, double maximum_range, uint32_t min_pass_through,
                                   double occupancy_threshold);
void kp_occupancy_destroy(kp_occupancy * g);
/* info = {width, height, width step}; offset = world position of cell (0,0) */
void kp_occupancy_info(const kp_occupancy * g, int32_t info[3], double offset[2]);
const uint8_t * kp_occupancy_cells(const kp_occupancy * g);
const uint32_t * kp_occupancy_pass(const kp_occupancy * g);
const uint32_t * kp_occupancy_hits(const kp_occupancy * g);

#ifdef __cplusplus
}
#endif
#endif
