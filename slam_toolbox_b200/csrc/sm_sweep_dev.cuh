// Block-level reductions and the response expression shared by the sweep kernels (sm_sweep.cu, sm_tile.cu).
#pragma once
#include "sm_math.cuh"
#include "sm_types.cuh"

namespace b200 {

__device__ __forceinline__ double block_max(double v, double * scratch)
{
  for (int o = 16; o > 0; o >>= 1) {
    double other = __shfl_xor_sync(0xffffffffu, v, o);
    v = other > v ? other : v;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  double r = scratch[0];
  for (int i = 1; i < nw; ++i) r = scratch[i] > r ? scratch[i] : r;
  __syncthreads();
  return r;
}

// exclusive prefix sum of one int per thread over the block; returns the exclusive value and
// writes the block total
__device__ __forceinline__ int block_exclusive_scan(int v, int * scratch, int & total)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();
  if (lane == 31) scratch[warp] = inc;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  int base = 0, tot = 0;
  for (int i = 0; i < nw; ++i) {
    if (i < warp) base += scratch[i];
    tot += scratch[i];
  }
  __syncthreads();
  total = tot;
  return base + inc - v;
}

// response of pose (xy, a) from its integer sum, exactly as ScanMatcher::operator() builds it
// (M.cpp:670-685)
__device__ __forceinline__ double pose_response(const SweepDev & d, int q, int sum, int x, int y, int a)
{
  double r = (double)sum;
  r /= d.norm;
  if (d.do_penalize && !double_equal(r, 0.0)) {
    double dp = distance_penalty(d.sqx[q * d.nX + x], d.sqy[q * d.nY + y], d.dist_var, d.min_dist_pen);
    double ap = d.angpen[q * d.nA + a];
    r *= (dp * ap);
  }
  return r;
}


}  // namespace b200
