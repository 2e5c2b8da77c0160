// Reference-side binding #1: karto::ScanSolver implemented on the b200slam C ABI.
// In a slam_toolbox workspace this class is exported with PLUGINLIB_EXPORT_CLASS and listed in
// solver_plugins.xml; here it is compiled against the reference headers (+ oracle/stubs for the absent
// Boost / Eigen / rclcpp headers) and driven by the reference's own karto::Mapper in integration/replay_driver.cpp.
#pragma once
#include <unordered_map>
#include <vector>

#include "karto_sdk/Mapper.h"
#include "b200slam.h"

namespace solver_plugins {

class B200Solver : public karto::ScanSolver
{
public:
  B200Solver() { b200pg_create(nullptr, &h_); }
  ~B200Solver() override { b200pg_destroy(h_); }

  void Configure(rclcpp_lifecycle::LifecycleNode::SharedPtr) override {}

  void Compute() override   // solvers/ceres_solver.cpp:214-269
  {
    ++computes_;
    if (getenv("B200_TRACE")) fprintf(stderr, "Compute\n");
    b200pg_summary s;
    if (b200pg_solve(h_, &s) != B200_OK) return;   // unusable: corrections untouched, like the reference
    solve_ms_ += s.solve_ms;
    const int n = b200pg_num_nodes(h_);
    std::vector<int32_t> ids(n);
    std::vector<double> p(3 * (size_t)n);
    const int m = b200pg_get_corrections(h_, ids.data(), p.data(), n);
    corrections_.clear();
    corrections_.reserve(m);
    for (int i = 0; i < m; ++i) corrections_.emplace_back(ids[i], karto::Pose2(p[3 * i], p[3 * i + 1], p[3 * i + 2]));
  }
  const karto::ScanSolver::IdPoseVector & GetCorrections() const override { return corrections_; }
  void Clear() override { corrections_.clear(); b200pg_clear(h_); }
  void Reset() override { corrections_.clear(); ids_.clear(); b200pg_reset(h_); }

  void AddNode(karto::Vertex<karto::LocalizedRangeScan> * v) override   // ceres_solver.cpp:317-336
  {
    if (!v) return;
    if (getenv("B200_TRACE")) fprintf(stderr, "AddNode %p\n", (void*)v);
    const karto::Pose2 p = v->GetObject()->GetCorrectedPose();
    const double pose[3] = {p.GetX(), p.GetY(), p.GetHeading()};
    if (b200pg_add_node(h_, v->GetObject()->GetUniqueId(), pose) == B200_OK) ids_.push_back(v->GetObject()->GetUniqueId());
  }
  void AddConstraint(karto::Edge<karto::LocalizedRangeScan> * e) override   // ceres_solver.cpp:339-392
  {
    if (!e) return;
    if (getenv("B200_TRACE")) fprintf(stderr, "AddConstraint %p label %p\n", (void*)e, (void*)e->GetLabel());
    karto::LinkInfo * li = static_cast<karto::LinkInfo *>(e->GetLabel());
    const karto::Pose2 d = li->GetPoseDifference();
    const karto::Matrix3 c = li->GetCovariance();
    const double z[3] = {d.GetX(), d.GetY(), d.GetHeading()};
    double cov[9];
    for (int r = 0; r < 3; ++r) for (int k = 0; k < 3; ++k) cov[3 * r + k] = c(r, k);
    b200pg_add_edge(h_, e->GetSource()->GetObject()->GetUniqueId(), e->GetTarget()->GetObject()->GetUniqueId(), z, cov);
  }
  void RemoveNode(kt_int32s id) override
  {
    if (b200pg_remove_node(h_, id) != B200_OK) return;
    for (size_t i = 0; i < ids_.size(); ++i)
      if (ids_[i] == id) { ids_.erase(ids_.begin() + i); break; }
  }
  void RemoveConstraint(kt_int32s a, kt_int32s b) override { b200pg_remove_edge(h_, a, b); }
  void ModifyNode(const int & id, Eigen::Vector3d pose) override
  {
    const double p[3] = {pose(0), pose(1), pose(2)};
    b200pg_modify_node(h_, id, p);
  }
  void GetNodeOrientation(const int & id, double & yaw) override
  {
    double p[3];
    if (b200pg_get_node(h_, id, p) == B200_OK) yaw = p[2];
  }

  // the raw node store for visualisation (ceres_solver.cpp:474-479, used by src/loop_closure_assistant.cpp:161):
  // a host copy of the nodes as the library holds them now, refreshed on every call
  std::unordered_map<int, Eigen::Vector3d> * getGraph() override
  {
    graph_.clear();
    for (int id : ids_) {
      double p[3];
      if (b200pg_get_node(h_, id, p) != B200_OK) continue;
      Eigen::Vector3d v;
      v(0) = p[0]; v(1) = p[1]; v(2) = p[2];
      graph_[id] = v;
    }
    return &graph_;
  }

  int computes() const { return computes_; }
  double solve_ms() const { return solve_ms_; }

private:
  b200pg * h_ = nullptr;
  karto::ScanSolver::IdPoseVector corrections_;
  std::vector<int> ids_;                                  // node ids in insertion order
  std::unordered_map<int, Eigen::Vector3d> graph_;
  int computes_ = 0;
  double solve_ms_ = 0.0;
};

}  // namespace solver_plugins
