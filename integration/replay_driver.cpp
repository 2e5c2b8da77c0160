// cfg3 replay driver: the reference's own karto::Mapper::Process (Mapper.cpp:2679-2749) fed with posed
// scans, with the GPU ScanSolver adapter installed through Mapper::SetScanSolver and -- when linked into
// libreplay_b200.so -- every MatchScan redirected to the GPU by scan_matcher_b200.cpp.
// libreplay_ref.so is the same driver linked WITHOUT the matcher shim (reference CPU matcher).
#include <chrono>
#include <csignal>
#include <cstdlib>
#include <execinfo.h>
#include <string>
#include <unistd.h>

#include "karto_sdk/Mapper.h"
#include "b200_solver.hpp"
#include "occupancy_b200.hpp"

using namespace karto;

namespace {
struct Replay {
  Mapper mapper;
  solver_plugins::B200Solver * solver = nullptr;
  std::vector<LocalizedRangeScan *> scans;
  double process_seconds = 0.0;
  int processed = 0;
};
const char * kLaser = "laser0";
}

static void segv_handler(int)
{
  void * frames[64];
  int n = backtrace(frames, 64);
  backtrace_symbols_fd(frames, n, 2);
  _exit(139);
}

extern "C" {

int krep_init_laser(double min_angle, double max_angle, double ang_res, double min_range, double max_range, double range_threshold)
{
  if (getenv("B200_TRACE")) signal(SIGSEGV, segv_handler);
  Name nm(kLaser);
  LaserRangeFinder * l = LaserRangeFinder::CreateLaserRangeFinder(LaserRangeFinder_Custom, nm);
  l->SetMinimumRange(min_range); l->SetMaximumRange(max_range);
  l->SetMinimumAngle(min_angle); l->SetMaximumAngle(max_angle);
  l->SetAngularResolution(ang_res); l->SetRangeThreshold(range_threshold);
  l->SetOffsetPose(Pose2(0.0, 0.0, 0.0));
  SensorManager::GetInstance()->RegisterSensor(l);
  return static_cast<int>(l->GetNumberOfRangeReadings());
}

void * krep_create(int use_solver)
{
  Replay * r = new Replay();
  if (use_solver) {
    r->solver = new solver_plugins::B200Solver();
    r->mapper.SetScanSolver(r->solver);
  }
  return r;
}

// B200Solver::ConfigureFromStrings (the mapping CeresSolver::Configure applies to the ceres_* ROS parameters): returns the
// number of keys applied, -1 without a solver
int krep_solver_configure(void * rp, const char * key, const char * value)
{
  Replay * r = static_cast<Replay *>(rp);
  if (!r->solver) return -1;
  return r->solver->ConfigureFromStrings({{std::string(key), std::string(value)}});
}

// Mapper::Reset (Mapper.cpp:2656-2677) deletes both scan matchers; krep_destroy deletes the mapper.  With the matcher shim
// linked, b200_shim_live_handles() must drop to 0 afterwards (no leaked device state).
void krep_reset_mapper(void * rp) { static_cast<Replay *>(rp)->mapper.Reset(); }
void krep_destroy(void * rp)
{
  Replay * r = static_cast<Replay *>(rp);
  solver_plugins::B200Solver * s = r->solver;
  delete r;
  delete s;
}

int krep_set(void * rp, const char * name, double v)
{
  Mapper * m = &static_cast<Replay *>(rp)->mapper;
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

// feeds one scan with its odometric pose; returns 1 if the mapper processed (kept) it
int krep_process(void * rp, const double * ranges, int n, const double odom[3], int id)
{
  Replay * r = static_cast<Replay *>(rp);
  RangeReadingsVector rr(ranges, ranges + n);
  LocalizedRangeScan * s = new LocalizedRangeScan(Name(kLaser), rr);
  Pose2 p(odom[0], odom[1], odom[2]);
  s->SetOdometricPose(p);
  s->SetCorrectedPose(p);
  s->SetTime(static_cast<double>(id));
  auto t0 = std::chrono::steady_clock::now();
  bool ok = false;
  try {
    ok = r->mapper.Process(s);
  } catch (const std::exception & e) {
    std::fprintf(stderr, "krep_process: %s\n", e.what());
  }
  r->process_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (ok) { r->scans.push_back(s); ++r->processed; } else { delete s; }
  return ok ? 1 : 0;
}

int krep_num_scans(void * rp) { return static_cast<int>(static_cast<Replay *>(rp)->scans.size()); }

// corrected poses (after all loop closures) of the processed scans, in processing order
void krep_poses(void * rp, double * out)
{
  Replay * r = static_cast<Replay *>(rp);
  for (size_t i = 0; i < r->scans.size(); ++i) {
    const Pose2 & p = r->scans[i]->GetCorrectedPose();
    out[3 * i] = p.GetX(); out[3 * i + 1] = p.GetY(); out[3 * i + 2] = p.GetHeading();
  }
}

// ScanSolver::getGraph() of the adapter: number of nodes; ids / poses filled up to cap
int krep_solver_graph(void * rp, int * ids, double * poses, int cap)
{
  Replay * r = static_cast<Replay *>(rp);
  if (!r->solver) return 0;
  std::unordered_map<int, Eigen::Vector3d> * g = r->solver->getGraph();
  int k = 0;
  for (const auto & kv : *g) {
    if (k < cap) { ids[k] = kv.first; poses[3 * k] = kv.second(0); poses[3 * k + 1] = kv.second(1); poses[3 * k + 2] = kv.second(2); }
    ++k;
  }
  return k;
}

// stats: process seconds, solver computes, solver device ms, graph edges, graph vertices
void krep_stats(void * rp, double out[5])
{
  Replay * r = static_cast<Replay *>(rp);
  out[0] = r->process_seconds;
  out[1] = r->solver ? r->solver->computes() : 0;
  out[2] = r->solver ? r->solver->solve_ms() : 0;
  out[3] = static_cast<double>(r->mapper.GetGraph() ? r->mapper.GetGraph()->GetEdges().size() : 0);
  out[4] = static_cast<double>(r->scans.size());
}

// map publish over all processed scans (SMapper::getOccupancyGrid, src/slam_mapper.cpp:63-69):
// use_gpu = 0 the reference's OccupancyGrid::CreateFromScans, 1 the b200og binding.  info = {width, height,
// width step}; cells (if not NULL, cap bytes) receives the grid bytes.  Returns wall seconds, < 0 on failure.
double krep_occupancy(void * rp, double resolution, int use_gpu, int info[3], double offset[2], unsigned char * cells, long cap)
{
  Replay * r = static_cast<Replay *>(rp);
  LocalizedRangeScanVector v(r->scans.begin(), r->scans.end());
  auto t0 = std::chrono::steady_clock::now();
  OccupancyGrid * g = use_gpu ? karto::b200::CreateOccupancyGridFromScans(v, resolution) : OccupancyGrid::CreateFromScans(v, resolution);
  const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (!g) return -1.0;
  info[0] = g->GetWidth(); info[1] = g->GetHeight(); info[2] = g->GetWidthStep();
  offset[0] = g->GetCoordinateConverter()->GetOffset().GetX();
  offset[1] = g->GetCoordinateConverter()->GetOffset().GetY();
  if (cells && cap >= static_cast<long>(g->GetDataSize())) std::memcpy(cells, g->GetDataPointer(), g->GetDataSize());
  delete g;
  return sec;
}

}  // extern "C"
