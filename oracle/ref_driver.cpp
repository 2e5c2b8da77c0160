// ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into the product library.
//
// Thin extern "C" driver around the UNMODIFIED reference karto_sdk
// (/root/reference/lib/karto_sdk/src/{Karto,Mapper}.cpp, compiled where they lie by
// oracle/Makefile into oracle/_ref/libkarto_ref.so).  It lets the Python tests and
// bench.py's cpu_baseline / --impl reference leg run the reference's own
// karto::ScanMatcher::MatchScan (Mapper.cpp:534-639) and read its correlation grid,
// lookup tables and intermediate results, so the restated C oracle (oracle/karto_port.c)
// and the CUDA path can be pinned against the reference itself.
//
// No reference source is copied here: this file only *calls* the reference API.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <vector>
#include <atomic>
#include <sstream>
#include <iostream>
#include <fstream>
#include <queue>
#include <set>
#include <list>
#include <unordered_map>
#include <chrono>
#include <shared_mutex>
#include <mutex>
#include <memory>
#include <algorithm>
#include <stdexcept>
#include <limits>
#include <iomanip>

// expose ScanMatcher internals (FindValidPoints, AddScans, GetResponse) to this TU only
#define private public
#define protected public
#include "karto_sdk/Mapper.h"
#undef private
#undef protected

using namespace karto;

namespace {
const char * kLaserName = "laser0";
LaserRangeFinder * g_laser = nullptr;
}

extern "C" {

// Registers the one laser all oracle scans use. Returns number of range readings.
int kref_init_laser(double min_angle, double max_angle, double ang_res, double min_range,
                    double max_range, double range_threshold)
{
  if (g_laser != nullptr) {
    return static_cast<int>(g_laser->GetNumberOfRangeReadings());
  }
  Name nm(kLaserName);
  LaserRangeFinder * l = LaserRangeFinder::CreateLaserRangeFinder(LaserRangeFinder_Custom, nm);
  l->SetMinimumRange(min_range);
  l->SetMaximumRange(max_range);
  l->SetMinimumAngle(min_angle);
  l->SetMaximumAngle(max_angle);
  l->SetAngularResolution(ang_res);
  l->SetRangeThreshold(range_threshold);
  l->SetOffsetPose(Pose2(0.0, 0.0, 0.0));
  SensorManager::GetInstance()->RegisterSensor(l);
  g_laser = l;
  return static_cast<int>(l->GetNumberOfRangeReadings());
}

void kref_laser_set_range_threshold(double rt) { if (g_laser) g_laser->SetRangeThreshold(rt); }

void * kref_mapper_create() { return new Mapper(); }
void kref_mapper_destroy(void * m) { delete static_cast<Mapper *>(m); }

// ROS-style parameter names (src/slam_mapper.cpp:96-358 -> Mapper::setParam*, Mapper.cpp:2452-2600)
int kref_mapper_set(void * mp, const char * name, double v)
{
  Mapper * m = static_cast<Mapper *>(mp);
  std::string n(name);
  if (n == "coarse_search_angle_offset") m->setParamCoarseSearchAngleOffset(v);
  else if (n == "coarse_angle_resolution") m->setParamCoarseAngleResolution(v);
  else if (n == "fine_search_angle_offset") m->setParamFineSearchAngleOffset(v);
  else if (n == "distance_variance_penalty") m->setParamDistanceVariancePenalty(v);
  else if (n == "angle_variance_penalty") m->setParamAngleVariancePenalty(v);
  else if (n == "minimum_distance_penalty") m->setParamMinimumDistancePenalty(v);
  else if (n == "minimum_angle_penalty") m->setParamMinimumAnglePenalty(v);
  else if (n == "use_response_expansion") m->setParamUseResponseExpansion(v != 0.0);
  else if (n == "correlation_search_space_dimension") m->setParamCorrelationSearchSpaceDimension(v);
  else if (n == "correlation_search_space_resolution") m->setParamCorrelationSearchSpaceResolution(v);
  else if (n == "correlation_search_space_smear_deviation") m->setParamCorrelationSearchSpaceSmearDeviation(v);
  else if (n == "loop_search_space_dimension") m->setParamLoopSearchSpaceDimension(v);
  else if (n == "loop_search_space_resolution") m->setParamLoopSearchSpaceResolution(v);
  else if (n == "loop_search_space_smear_deviation") m->setParamLoopSearchSpaceSmearDeviation(v);
  else if (n == "minimum_travel_distance") m->setParamMinimumTravelDistance(v);
  else if (n == "minimum_travel_heading") m->setParamMinimumTravelHeading(v);
  else if (n == "scan_buffer_size") m->setParamScanBufferSize(static_cast<int>(v));
  else if (n == "scan_buffer_maximum_scan_distance") m->setParamScanBufferMaximumScanDistance(v);
  else if (n == "link_match_minimum_response_fine") m->setParamLinkMatchMinimumResponseFine(v);
  else if (n == "link_scan_maximum_distance") m->setParamLinkScanMaximumDistance(v);
  else if (n == "loop_search_maximum_distance") m->setParamLoopSearchMaximumDistance(v);
  else if (n == "do_loop_closing") m->setParamDoLoopClosing(v != 0.0);
  else if (n == "loop_match_minimum_chain_size") m->setParamLoopMatchMinimumChainSize(static_cast<int>(v));
  else if (n == "loop_match_maximum_variance_coarse") m->setParamLoopMatchMaximumVarianceCoarse(v);
  else if (n == "loop_match_minimum_response_coarse") m->setParamLoopMatchMinimumResponseCoarse(v);
  else if (n == "loop_match_minimum_response_fine") m->setParamLoopMatchMinimumResponseFine(v);
  else if (n == "use_scan_matching") m->setParamUseScanMatching(v != 0.0);
  else if (n == "use_scan_barycenter") m->setParamUseScanBarycenter(v != 0.0);
  else if (n == "minimum_time_interval") m->setParamMinimumTimeInterval(v);
  else return -1;
  return 0;
}

// ScanMatcher::Create (Mapper.cpp:477-522). Returns NULL on invalid params like the reference.
void * kref_matcher_create(void * mapper, double search_size, double resolution, double smear,
                           double range_threshold)
{
  try {
    return ScanMatcher::Create(static_cast<Mapper *>(mapper), search_size, resolution, smear,
                               range_threshold);
  } catch (const std::exception & e) {
    std::cerr << "kref_matcher_create: " << e.what() << std::endl;
    return nullptr;
  }
}
void kref_matcher_destroy(void * sm) { delete static_cast<ScanMatcher *>(sm); }

void * kref_scan_create(const double * ranges, int n, const double pose[3], int unique_id)
{
  RangeReadingsVector r(ranges, ranges + n);
  LocalizedRangeScan * s = new LocalizedRangeScan(Name(kLaserName), r);
  Pose2 p(pose[0], pose[1], pose[2]);
  s->SetOdometricPose(p);
  s->SetCorrectedPose(p);
  s->SetUniqueId(unique_id);
  s->SetStateId(unique_id);
  s->SetTime(static_cast<double>(unique_id));
  (void)s->GetPointReadings();   // force Update() now so later concurrent reads are clean
  return s;
}
void kref_scan_destroy(void * s) { delete static_cast<LocalizedRangeScan *>(s); }

void kref_scan_set_pose(void * sp, const double pose[3])
{
  static_cast<LocalizedRangeScan *>(sp)->SetCorrectedPoseAndUpdate(Pose2(pose[0], pose[1], pose[2]));
}

// unfiltered point readings (Karto.h:5613-5628, 5644-5704), interleaved x,y; returns count
int kref_scan_points(void * sp, double * xy, int cap)
{
  const PointVectorDouble & pts = static_cast<LocalizedRangeScan *>(sp)->GetPointReadings(false);
  int n = static_cast<int>(pts.size());
  for (int i = 0; i < n && i < cap; ++i) { xy[2 * i] = pts[i].GetX(); xy[2 * i + 1] = pts[i].GetY(); }
  return n;
}
void kref_scan_sensor_pose(void * sp, double out[3])
{
  Pose2 p = static_cast<LocalizedRangeScan *>(sp)->GetSensorPose();
  out[0] = p.GetX(); out[1] = p.GetY(); out[2] = p.GetHeading();
}

static void put(const Pose2 & p, const Matrix3 & c, double mean[3], double cov[9])
{
  mean[0] = p.GetX(); mean[1] = p.GetY(); mean[2] = p.GetHeading();
  for (int r = 0; r < 3; ++r) for (int k = 0; k < 3; ++k) cov[3 * r + k] = c(r, k);
}

// karto::ScanMatcher::MatchScan<LocalizedRangeScanVector> (Mapper.cpp:534-639)
double kref_match(void * smp, void * query, void ** base, int nbase, int do_penalize, int do_refine,
                  double mean[3], double cov[9])
{
  ScanMatcher * sm = static_cast<ScanMatcher *>(smp);
  LocalizedRangeScanVector v;
  for (int i = 0; i < nbase; ++i) v.push_back(static_cast<LocalizedRangeScan *>(base[i]));
  Pose2 m; Matrix3 c;
  double r = sm->MatchScan(static_cast<LocalizedRangeScan *>(query), v, m, c, do_penalize != 0,
                           do_refine != 0);
  put(m, c, mean, cov);
  return r;
}

// Only the raster step of MatchScan: centre the grid on the query (Mapper.cpp:560-569) and AddScans (:574).
void kref_raster(void * smp, void * query, void ** base, int nbase)
{
  ScanMatcher * sm = static_cast<ScanMatcher *>(smp);
  LocalizedRangeScan * q = static_cast<LocalizedRangeScan *>(query);
  Pose2 scanPose = q->GetSensorPose();
  Rectangle2<kt_int32s> roi = sm->m_pCorrelationGrid->GetROI();
  Vector2<kt_double> offset;
  offset.SetX(scanPose.GetX() - (0.5 * (roi.GetWidth() - 1) * sm->m_pCorrelationGrid->GetResolution()));
  offset.SetY(scanPose.GetY() - (0.5 * (roi.GetHeight() - 1) * sm->m_pCorrelationGrid->GetResolution()));
  sm->m_pCorrelationGrid->GetCoordinateConverter()->SetOffset(offset);
  LocalizedRangeScanVector v;
  for (int i = 0; i < nbase; ++i) v.push_back(static_cast<LocalizedRangeScan *>(base[i]));
  sm->AddScans(v, scanPose.GetPosition());
}

// public CorrelateScan (Mapper.cpp:712-862) on the grid as last rasterised
double kref_correlate(void * smp, void * query, const double center[3], const double sp_off[2],
                      const double sp_res[2], double ang_off, double ang_res, int do_penalize,
                      int fine, double mean[3], double cov_inout[9])
{
  ScanMatcher * sm = static_cast<ScanMatcher *>(smp);
  Pose2 m; Matrix3 c;
  for (int r = 0; r < 3; ++r) for (int k = 0; k < 3; ++k) c(r, k) = cov_inout[3 * r + k];
  double resp = sm->CorrelateScan(static_cast<LocalizedRangeScan *>(query),
      Pose2(center[0], center[1], center[2]), Vector2<kt_double>(sp_off[0], sp_off[1]),
      Vector2<kt_double>(sp_res[0], sp_res[1]), ang_off, ang_res, do_penalize != 0, m, c, fine != 0);
  put(m, c, mean, cov_inout);
  return resp;
}

// info: width,height,stride,roi_x,roi_y,roi_w,roi_h,data_size,kernel_size ; off = coordinate offset
void kref_grid_info(void * smp, int info[9], double off[2], double * scale)
{
  CorrelationGrid * g = static_cast<ScanMatcher *>(smp)->GetCorrelationGrid();
  info[0] = g->GetWidth(); info[1] = g->GetHeight(); info[2] = g->GetWidthStep();
  info[3] = g->GetROI().GetX(); info[4] = g->GetROI().GetY();
  info[5] = g->GetROI().GetWidth(); info[6] = g->GetROI().GetHeight();
  info[7] = g->GetDataSize(); info[8] = g->m_KernelSize;
  off[0] = g->GetCoordinateConverter()->GetOffset().GetX();
  off[1] = g->GetCoordinateConverter()->GetOffset().GetY();
  *scale = g->GetCoordinateConverter()->GetScale();
}
void kref_grid_copy(void * smp, uint8_t * out)
{
  CorrelationGrid * g = static_cast<ScanMatcher *>(smp)->GetCorrelationGrid();
  std::memcpy(out, g->GetDataPointer(), g->GetDataSize());
}
void kref_kernel_copy(void * smp, uint8_t * out)
{
  CorrelationGrid * g = static_cast<ScanMatcher *>(smp)->GetCorrelationGrid();
  std::memcpy(out, g->m_pKernel, g->m_KernelSize * g->m_KernelSize);
}

// GridIndexLookup::ComputeOffsets (Karto.h:6797-6894) on a fresh lookup bound to the matcher's
// grid (uses the grid's current coordinate offset). out = nAngles x nReadings int32. Returns nAngles.
int kref_offsets(void * smp, void * query, double angle_center, double angle_offset,
                 double angle_res, int32_t * out, int cap_per_angle)
{
  ScanMatcher * sm = static_cast<ScanMatcher *>(smp);
  GridIndexLookup<kt_int8u> lut(sm->GetCorrelationGrid());
  lut.ComputeOffsets(static_cast<LocalizedRangeScan *>(query), angle_center, angle_offset, angle_res);
  int n_angles = static_cast<int>(math::Round(angle_offset * 2.0 / angle_res) + 1);
  for (int a = 0; a < n_angles; ++a) {
    const LookupArray * la = lut.GetLookupArray(a);
    int n = static_cast<int>(la->GetSize());
    for (int i = 0; i < n && i < cap_per_angle; ++i) out[a * cap_per_angle + i] = la->GetArrayPointer()[i];
  }
  return n_angles;
}

// ScanMatcher::FindValidPoints (Mapper.cpp:1113-1164); returns count, writes interleaved xy
int kref_find_valid_points(void * smp, void * scan, const double viewpoint[2], double * xy, int cap)
{
  ScanMatcher * sm = static_cast<ScanMatcher *>(smp);
  PointVectorDouble v = sm->FindValidPoints(static_cast<LocalizedRangeScan *>(scan),
      Vector2<kt_double>(viewpoint[0], viewpoint[1]));
  int n = static_cast<int>(v.size());
  for (int i = 0; i < n && i < cap; ++i) { xy[2 * i] = v[i].GetX(); xy[2 * i + 1] = v[i].GetY(); }
  return n;
}

// Loop-closure style sweep on host threads: candidate chain j = scans[chain_start[j] .. chain_start[j+1]).
// One reference ScanMatcher per thread (the class is not re-entrant, Mapper.h:1496-1503).
// out: resp[nchains], mean[3*nchains], cov[9*nchains]. Returns wall seconds.
double kref_sweep(void ** matchers, int nthreads, void * query, void ** scans, const int * chain_start,
                  int nchains, int do_penalize, int do_refine, double * resp, double * mean, double * cov)
{
  auto t0 = std::chrono::steady_clock::now();
  std::atomic<int> next(0);
  auto work = [&](int t) {
    ScanMatcher * sm = static_cast<ScanMatcher *>(matchers[t]);
    for (;;) {
      int j = next.fetch_add(1);
      if (j >= nchains) break;
      LocalizedRangeScanVector v;
      for (int i = chain_start[j]; i < chain_start[j + 1]; ++i) v.push_back(static_cast<LocalizedRangeScan *>(scans[i]));
      Pose2 m; Matrix3 c;
      resp[j] = sm->MatchScan(static_cast<LocalizedRangeScan *>(query), v, m, c, do_penalize != 0, do_refine != 0);
      put(m, c, mean + 3 * j, cov + 9 * j);
    }
  };
  std::vector<std::thread> ts;
  for (int t = 1; t < nthreads; ++t) ts.emplace_back(work, t);
  work(0);
  for (auto & t : ts) t.join();
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// LinkInfo::Update (Mapper.h:174-188): edge measurement + rotated covariance
void kref_link_info(const double p1[3], const double p2[3], const double cov[9], double diff[3], double cov_out[9])
{
  LinkInfo li(Pose2(p1[0], p1[1], p1[2]), Pose2(p2[0], p2[1], p2[2]), [&] {
      Matrix3 c; for (int r = 0; r < 3; ++r) for (int k = 0; k < 3; ++k) c(r, k) = cov[3 * r + k]; return c; }());
  put(li.GetPoseDifference(), li.GetCovariance(), diff, cov_out);
}

// Matrix3::Inverse (Karto.h:2533-2577), used by CeresSolver::AddConstraint (solvers/ceres_solver.cpp:364-376)
void kref_matrix3_inverse(const double m[9], double out[9])
{
  Matrix3 c; for (int r = 0; r < 3; ++r) for (int k = 0; k < 3; ++k) c(r, k) = m[3 * r + k];
  Matrix3 inv = c.Inverse();
  for (int r = 0; r < 3; ++r) for (int k = 0; k < 3; ++k) out[3 * r + k] = inv(r, k);
}

// ---- occupancy grid (Karto.h:5883-6330) ----------------------------------------------------
// karto::OccupancyGrid::CreateFromScans (Karto.h:5946-5961) as slam_toolbox calls it
// (src/slam_mapper.cpp:63-69).  min_pass_through < 0 and occupancy_threshold < 0 take the static
// entry point untouched (defaults 2 and 0.1, Karto.h:5921-5922); otherwise the same three steps are
// issued here with the two parameters set before the counters are evaluated.
void * kref_occupancy_create(void ** scans, int n, double resolution, int min_pass_through, double occupancy_threshold)
{
  LocalizedRangeScanVector v;
  for (int i = 0; i < n; ++i) v.push_back(static_cast<LocalizedRangeScan *>(scans[i]));
  if (min_pass_through < 0 && occupancy_threshold < 0) return OccupancyGrid::CreateFromScans(v, resolution);
  if (v.empty()) return nullptr;
  kt_int32s w, h; Vector2<kt_double> off;
  OccupancyGrid::ComputeDimensions(v, resolution, w, h, off);
  OccupancyGrid * g = new OccupancyGrid(w, h, off, resolution);
  if (min_pass_through >= 0) g->SetMinPassThrough(static_cast<kt_int32u>(min_pass_through));
  if (occupancy_threshold >= 0) g->SetOccupancyThreshold(occupancy_threshold);
  g->CreateFromScans(v);
  return g;
}
void kref_occupancy_destroy(void * g) { delete static_cast<OccupancyGrid *>(g); }
// info = {width, height, width step}; offset = world position of cell (0,0)
void kref_occupancy_info(void * gp, int info[3], double offset[2])
{
  OccupancyGrid * g = static_cast<OccupancyGrid *>(gp);
  info[0] = g->GetWidth(); info[1] = g->GetHeight(); info[2] = g->GetWidthStep();
  offset[0] = g->GetCoordinateConverter()->GetOffset().GetX();
  offset[1] = g->GetCoordinateConverter()->GetOffset().GetY();
}
// cells (width step x height bytes) and, if not NULL, the pass / hit counters (same layout)
void kref_occupancy_copy(void * gp, uint8_t * cells, uint32_t * pass, uint32_t * hits)
{
  OccupancyGrid * g = static_cast<OccupancyGrid *>(gp);
  const size_t n = static_cast<size_t>(g->GetDataSize());
  if (cells) std::memcpy(cells, g->GetDataPointer(), n);
  if (pass) std::memcpy(pass, g->GetCellPassCounts()->GetDataPointer(), n * sizeof(uint32_t));
  if (hits) std::memcpy(hits, g->GetCellHitsCounts()->GetDataPointer(), n * sizeof(uint32_t));
}

}  // extern "C"
