// Exact-arithmetic helpers shared by the host orchestration and the device kernels of the
// scan matcher.  Every function is __host__ __device__ and uses only IEEE-754 +,-,*,/ and
// comparisons in the reference's operand order, so host (g++ -ffp-contract=off) and device
// (nvcc -fmad=false) produce the same bits as the reference's x86-64 build.  libm calls
// (sin, cos, atan2, exp, pow, hypot) are host-only: the device never evaluates them on the
// parity path (hard part 1).
//
// Reference lines: Math.h = lib/karto_sdk/include/karto_sdk/Math.h, K.h = .../Karto.h,
// M.cpp = lib/karto_sdk/src/Mapper.cpp.
#pragma once
#include <cstdint>
#include <cmath>

#ifdef __CUDACC__
#define B200_HD __host__ __device__ __forceinline__
#else
#define B200_HD inline
#endif

namespace b200 {

constexpr double kPi = 3.14159265358979323846;        // Math.h:32
constexpr double k2Pi = 6.28318530717958647692;       // Math.h:33
constexpr double kPi180 = 0.01745329251994329577;     // Math.h:35
constexpr double kTolerance = 1e-06;                  // Math.h:41
constexpr int32_t kInvalidScan = 2147483647;          // Math.h:47
constexpr double kMaxVariance = 500.0;                // M.cpp:52
constexpr double kDistancePenaltyGain = 0.2;          // M.cpp:53
constexpr double kAnglePenaltyGain = 0.2;             // M.cpp:54
constexpr int kOccupied = 100;                        // K.h:4380

B200_HD double round_half_away(double v)              // Math.h:87-90
{
  return v >= 0.0 ? floor(v + 0.5) : ceil(v - 0.5);
}
B200_HD bool double_equal(double a, double b)         // Math.h:136-140
{
  double delta = a - b;
  return delta < 0.0 ? delta >= -kTolerance : delta <= kTolerance;
}
B200_HD double square(double v) { return v * v; }
B200_HD double maximum(double a, double b) { return a > b ? a : b; }   // Math.h:112
// static_cast<kt_int32s>(double) as the reference's x86-64 build performs it (cvttsd2si):
// NaN, Inf and out-of-range values become INT32_MIN, which then fails every IsUpTo test.
B200_HD int32_t to_int32(double r)
{
  if (!(r > -2147483649.0 && r < 2147483648.0)) return INT32_MIN;
  return static_cast<int32_t>(r);
}
// CoordinateConverter::WorldToGrid, one axis (K.h:4421-4436)
B200_HD int32_t world_to_grid(double w, double offset, double scale)
{
  return to_int32(round_half_away((w - offset) * scale));
}
B200_HD bool is_up_to(int32_t v, int32_t m) { return v >= 0 && v < m; }   // Math.h:149

// One step of ScanMatcher::FindValidPoints' state machine (M.cpp:1113-1164).
struct ValidPointState {
  double fx, fy;      // firstPoint
  int32_t trailing;   // trailingPointIter
  bool first_time;
  B200_HD void init() { fx = 0.0; fy = 0.0; trailing = 0; first_time = true; }
  // Feeds point i; returns the half-open range [lo, hi) of point indices that become valid.
  B200_HD void step(int32_t i, double cx, double cy, double vx, double vy, int32_t & lo, int32_t & hi)
  {
    lo = hi = 0;
    if (first_time && !(cx != cx) && !(cy != cy)) { fx = cx; fy = cy; first_time = false; }
    double dx = fx - cx, dy = fy - cy;
    if (dx * dx + dy * dy > 0.1 * 0.1) {
      double a = vy - fy;
      double b = fx - vx;
      double c = fy * vx - fx * vy;
      double ss = cx * a + cy * b + c;
      fx = cx; fy = cy;
      if (ss < 0.0) {
        trailing = i;
      } else {
        lo = trailing; hi = i; trailing = i;
      }
    }
  }
};

// distance / angle penalty of ScanMatcher::operator() (M.cpp:671-685)
B200_HD double distance_penalty(double squareX, double squareY, double var, double min_pen)
{
  double squaredDistance = squareX + squareY;
  double p = 1.0 - (kDistancePenaltyGain * squaredDistance / var);
  return maximum(p, min_pen);
}
B200_HD double angle_penalty(double angle, double center_heading, double var, double min_pen)
{
  double sq = square(angle - center_heading);
  double p = 1.0 - (kAnglePenaltyGain * sq / var);
  return maximum(p, min_pen);
}

}  // namespace b200
