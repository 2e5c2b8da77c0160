// Oracle build shim: tbb::parallel_for_each over the search rows (Mapper.cpp:773).
// ORACLE_TBB_THREADS (env) > 1 runs rows on std::threads (rows write disjoint
// slots of m_pPoseResponse, Mapper.cpp:688-691); default is a serial loop.
#pragma once
#include <cstdlib>
#include <thread>
#include <vector>
#include <atomic>
namespace tbb {
inline int shim_threads() {
  static int n = [] { const char * e = std::getenv("ORACLE_TBB_THREADS"); int v = e ? std::atoi(e) : 1; return v < 1 ? 1 : v; }();
  return n;
}
template <class C, class F>
inline void parallel_for_each(C & c, const F & f) {
  int nt = shim_threads();
  if (nt <= 1 || c.size() < 2) { for (auto & v : c) f(v); return; }
  std::atomic<size_t> next(0);
  auto work = [&]() { for (;;) { size_t i = next.fetch_add(1); if (i >= c.size()) break; f(c[i]); } };
  std::vector<std::thread> ts;
  for (int t = 1; t < nt; ++t) ts.emplace_back(work);
  work();
  for (auto & t : ts) t.join();
}
}  // namespace tbb
