// Reference-side binding #3: karto::OccupancyGrid::CreateFromScans
// (Karto.h:5946-5961) through the b200slam C ABI.
//
// CreateFromScans is a static INLINE member of a header class, so there is no symbol to interpose: the
// binding is a drop-in function with the same signature and result type that a maintainer calls from the
// one place slam_toolbox builds its map, SMapper::getOccupancyGrid (src/slam_mapper.cpp:63-69):
//
//     return karto::b200::CreateOccupancyGridFromScans(mapper_->GetAllProcessedScans(), resolution);
//
// The returned karto::OccupancyGrid has the reference's width / height / offset / resolution and cell
// bytes, which is everything its consumers read (vis_utils::toNavMap, include/slam_toolbox/
// visualization_utils.hpp:108-146; map_saver; RayCast / IsFree).  The hit / pass counter grids inside it
// stay empty: they are scratch of the CPU build (fetch them with b200og_fetch when wanted).
#pragma once
#include <cstdio>
#include <cstring>
#include <vector>

#include "karto_sdk/Karto.h"
#include "b200slam.h"

namespace karto {
namespace b200 {

inline OccupancyGrid * CreateOccupancyGridFromScans(const LocalizedRangeScanVector & rScans, kt_double resolution,
                                                     float * device_ms = nullptr)
{
  std::vector<b200_scan> scans;
  LaserRangeFinder * laser = nullptr;
  static_assert(sizeof(Vector2<kt_double>) == 2 * sizeof(double), "Vector2<double> must be {x, y}");
  for (LocalizedRangeScan * s : rScans) {
    if (s == nullptr) continue;                                   // Karto.h:6094, :6123
    if (!laser) laser = s->GetLaserRangeFinder();
    b200_scan o;
    const PointVectorDouble & pts = s->GetPointReadings(false);  // Karto.h:6149
    o.n = static_cast<int32_t>(pts.size());
    o.ranges = s->GetRangeReadings();
    o.points_xy = reinterpret_cast<const double *>(pts.data());
    const Pose2 p = s->GetSensorPose();
    o.sensor_pose[0] = p.GetX(); o.sensor_pose[1] = p.GetY(); o.sensor_pose[2] = p.GetHeading();
    scans.push_back(o);
  }
  if (rScans.empty()) return NULL;                                // Karto.h:5950-5952
  b200og_params p;
  b200og_default_params(&p);                                      // MinPassThrough 2, OccupancyThreshold 0.1
  p.resolution = resolution;
  if (laser) {
    p.range_threshold = laser->GetRangeThreshold();
    p.minimum_range = laser->GetMinimumRange();
    p.maximum_range = laser->GetMaximumRange();
  }
  b200og * h = nullptr;
  b200og_info info;
  if (b200og_create_from_scans(&p, scans.data(), static_cast<int32_t>(scans.size()), &info, &h) != B200_OK) {
    std::fprintf(stderr, "OccupancyGrid::CreateFromScans (b200): %s\n", b200_last_error());
    return NULL;
  }
  OccupancyGrid * g = new OccupancyGrid(info.width, info.height, Vector2<kt_double>(info.offset[0], info.offset[1]), resolution);
  if (g->GetWidthStep() != info.stride || b200og_fetch(h, g->GetDataPointer(), nullptr, nullptr) != B200_OK) {
    std::fprintf(stderr, "OccupancyGrid::CreateFromScans (b200): %s\n", b200_last_error());
    delete g;
    g = NULL;
  }
  if (device_ms) b200og_kernel_ms(h, device_ms);
  b200og_destroy(h);
  return g;
}

}  // namespace b200
}  // namespace karto
