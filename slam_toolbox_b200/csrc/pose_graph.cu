// b200slam SE(2) pose-graph solver: the karto::ScanSolver surface
// (lib/karto_sdk/include/karto_sdk/Mapper.h:954-1065) as solver_plugins::CeresSolver
// implements it (solvers/ceres_solver.cpp, solvers/ceres_utils.h), rebuilt for one B200:
//
//   * problem: nodes (x, y, theta), edges with relative-pose measurement z and sqrt-information
//     U = chol(cov^-1).matrixU() (ceres_solver.cpp:364-376); residual r = U [R(th_a)^T (p_b - p_a)
//     - t ; wrap(th_b - th_a - th_ab)] (ceres_utils.h:84-100); first node constant (:228-241).
//   * outer loop: Ceres' trust-region Levenberg-Marquardt schedule (TrustRegionMinimizer /
//     LevenbergMarquardtStrategy / TrustRegionStepEvaluator) with the reference's options
//     (ceres_solver.cpp:158-186), driven from the host with one small D2H of scalars per step.
//   * inner solve: instead of SPARSE_NORMAL_CHOLESKY, block-Jacobi preconditioned CG on the
//     3x3-block normal equations, run to a tight tolerance inside ONE persistent cooperative
//     kernel per LM iteration (2 grid barriers per CG iteration, deterministic reductions).
//   * one fused kernel evaluates every edge's residual, both Jacobian blocks (analytic) and the
//     off-diagonal normal-equation block; a second gathers per-node diagonal blocks/gradients
//     through a CSR node->edge adjacency (no atomics: bit-reproducible results).
//
// Data layout (HBM, FP64): nodes AoS [N][3]; edges SoA-of-small-arrays: idx[E][2], z[E][3],
// U[E][6] (upper triangle), lin[E][30] = r(3) | A~(9) | B~(9) | M = A~^T B~ (9); node blocks
// Hd[N][6] (symmetric), g[N][3]. The whole 10k/40k problem is ~15 MB: L2 resident.
#include <cooperative_groups.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace b200 {

constexpr int kPgThreads = 256;
constexpr int kLin = 30;   // doubles per edge in the linearisation record
constexpr int kMaxPartials = 1024;

struct PgDev {
  int N, E;
  const int32_t * eidx;     // [E][2] node indices
  const double * z;         // [E][3]
  const double * U;         // [E][6] u00 u01 u02 u11 u12 u22
  const uint8_t * is_free;  // [N] 1 = optimised, 0 = constant / not in the problem
  const int32_t * adj_start;   // [N+1]
  const int32_t * adj;         // [2E] (edge << 1) | side   (side 0: node is a, 1: node is b)
  double * x;               // [N][3] current iterate
  double * xc;              // [N][3] candidate
  double * scale;           // [N][3] Jacobi column scaling
  double * lin;             // [E][30]
  double * Hd;              // [N][6] diag blocks of J~^T J~ (xx xy xt yy yt tt)
  double * g;               // [N][3] J~^T r
  double * diag;            // [N][3] LM diagonal (clamped squared column norms)
  double * y;               // [N][3] PCG solution of (H + D^2) y = g
  double * pr, * pz, * pp0, * pp1, * pq, * Minv;   // PCG work vectors [N][3], Minv [N][6]
  double * partial;         // [kMaxPartials * 4] per-CTA partial sums
  double * scalars;         // small result block
  int loss;                 // 0 none, 1 Huber, 2 Cauchy (ceres_solver.cpp:82-94)
  double loss_a;            // loss scale
};

__device__ __forceinline__ double wrap_angle(double a)   // ceres_utils.h:27-32
{
  const double two_pi = 2.0 * M_PI;
  return a - two_pi * floor((a + M_PI) / two_pi);
}

__device__ __forceinline__ double warp_sum(double v)
{
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v)
{
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// deterministic block reduction of up to 4 values; result valid in thread 0
template <int K>
__device__ __forceinline__ void block_sum(double (&v)[K], double * smem /* K*32 */)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) v[k] = warp_sum(v[k]);
  __syncthreads();
  if (lane == 0)
    for (int k = 0; k < K; ++k) smem[k * 32 + warp] = v[k];
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 0; k < K; ++k) {
      double s = 0;
      for (int w = 0; w < nw; ++w) s += smem[k * 32 + w];
      v[k] = s;
    }
  }
}

// residual of one edge at poses (pa, pb): PoseGraph2dErrorTerm::operator() (ceres_utils.h:84-100)
__device__ __forceinline__ void edge_residual(const double * pa, const double * pb, const double * z, const double * U,
                                              double & c, double & s, double & dx, double & dy, double r[3])
{
  sincos(pa[2], &s, &c);
  dx = pb[0] - pa[0]; dy = pb[1] - pa[1];
  const double e0 = c * dx + s * dy - z[0];
  const double e1 = -s * dx + c * dy - z[1];
  const double e2 = wrap_angle((pb[2] - pa[2]) - z[2]);
  r[0] = U[0] * e0 + U[1] * e1 + U[2] * e2;
  r[1] = U[3] * e1 + U[4] * e2;
  r[2] = U[5] * e2;
}

// ceres::LossFunction::Evaluate for the two losses the reference offers: rho(s) and rho'(s), s = ||r||^2.
// Both have rho'' <= 0, for which Ceres' Corrector reduces to scaling residual and Jacobian by sqrt(rho').
__device__ __forceinline__ void loss_eval(int loss, double a, double s, double & rho, double & rho1)
{
  const double b = a * a;
  if (loss == 1) {          // HuberLoss
    if (s > b) { const double r = sqrt(s); rho = 2.0 * a * r - b; rho1 = fmax(2.2250738585072014e-308, a / r); }
    else { rho = s; rho1 = 1.0; }
  } else if (loss == 2) {   // CauchyLoss
    const double sum = 1.0 + s / b;
    rho = b * log(sum); rho1 = fmax(2.2250738585072014e-308, 1.0 / sum);
  } else { rho = s; rho1 = 1.0; }
}

// Fused linearisation: residual, Jacobian blocks w.r.t. node a and b (analytic form of the
// reference's autodiff), Jacobi column scaling, off-diagonal normal block M = A~^T B~, cost partial.
// mode 0: full linearisation at d.x into d.lin ; mode 1: cost only at d.xc.
__global__ void __launch_bounds__(kPgThreads) k_pg_linearize(PgDev d, int mode)
{
  __shared__ double red[32];
  double cost[1] = {0.0};
  const double * X = mode == 0 ? d.x : d.xc;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < d.E; e += gridDim.x * blockDim.x) {
    const int a = d.eidx[2 * e], b = d.eidx[2 * e + 1];
    const double * pa = X + 3 * a, * pb = X + 3 * b;
    const double * U = d.U + 6 * e;
    double c, s, dx, dy, r[3];
    edge_residual(pa, pb, d.z + 3 * e, U, c, s, dx, dy, r);
    const double sq = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
    double w = 1.0;   // sqrt(rho'): Corrector::CorrectResiduals / CorrectJacobian for rho'' <= 0
    if (d.loss) {
      double rho, rho1;
      loss_eval(d.loss, d.loss_a, sq, rho, rho1);
      cost[0] += rho;
      w = sqrt(rho1);
    } else {
      cost[0] += sq;
    }
    if (mode == 1) continue;
    r[0] *= w; r[1] *= w; r[2] *= w;
    // de/d(xa,ya,tha) and de/d(xb,yb,thb)
    const double Ae[9] = {-c, -s, -s * dx + c * dy, s, -c, -c * dx - s * dy, 0.0, 0.0, -1.0};
    const double Be[9] = {c, s, 0.0, -s, c, 0.0, 0.0, 0.0, 1.0};
    double A[9], B[9];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      A[0 + j] = U[0] * Ae[0 + j] + U[1] * Ae[3 + j] + U[2] * Ae[6 + j];
      A[3 + j] = U[3] * Ae[3 + j] + U[4] * Ae[6 + j];
      A[6 + j] = U[5] * Ae[6 + j];
      B[0 + j] = U[0] * Be[0 + j] + U[1] * Be[3 + j] + U[2] * Be[6 + j];
      B[3 + j] = U[3] * Be[3 + j] + U[4] * Be[6 + j];
      B[6 + j] = U[5] * Be[6 + j];
    }
    const double fa = d.is_free[a] ? w : 0.0, fb = d.is_free[b] ? w : 0.0;
    const double * sa = d.scale + 3 * a, * sb = d.scale + 3 * b;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) { A[3 * i + j] *= fa * sa[j]; B[3 * i + j] *= fb * sb[j]; }
    double * L = d.lin + (size_t)kLin * e;
    L[0] = r[0]; L[1] = r[1]; L[2] = r[2];
#pragma unroll
    for (int k = 0; k < 9; ++k) { L[3 + k] = A[k]; L[12 + k] = B[k]; }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) L[21 + 3 * i + j] = A[i] * B[j] + A[3 + i] * B[3 + j] + A[6 + i] * B[6 + j];
  }
  block_sum<1>(cost, red);
  if (threadIdx.x == 0) d.partial[blockIdx.x] = 0.5 * cost[0];
}

// Per node: diagonal block and gradient of the (scaled) normal equations gathered over incident
// edges in CSR order; squared column norms; ||x - Plus(x, -g_unscaled)||_inf partial (Ceres'
// gradient_max_norm); ||x||^2 partial over the free parameters.
__global__ void __launch_bounds__(kPgThreads) k_pg_assemble(PgDev d, int first /* compute Jacobi scaling */)
{
  __shared__ double red[64];
  double gmax = 0.0;
  double sums[1] = {0.0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < d.N; i += gridDim.x * blockDim.x) {
    double h[6] = {0, 0, 0, 0, 0, 0}, g[3] = {0, 0, 0};
    for (int k = d.adj_start[i]; k < d.adj_start[i + 1]; ++k) {
      const int e = d.adj[k] >> 1, side = d.adj[k] & 1;
      const double * L = d.lin + (size_t)kLin * e;
      const double * J = L + (side ? 12 : 3);
      h[0] += J[0] * J[0] + J[3] * J[3] + J[6] * J[6];
      h[1] += J[0] * J[1] + J[3] * J[4] + J[6] * J[7];
      h[2] += J[0] * J[2] + J[3] * J[5] + J[6] * J[8];
      h[3] += J[1] * J[1] + J[4] * J[4] + J[7] * J[7];
      h[4] += J[1] * J[2] + J[4] * J[5] + J[7] * J[8];
      h[5] += J[2] * J[2] + J[5] * J[5] + J[8] * J[8];
      g[0] += J[0] * L[0] + J[3] * L[1] + J[6] * L[2];
      g[1] += J[1] * L[0] + J[4] * L[1] + J[7] * L[2];
      g[2] += J[2] * L[0] + J[5] * L[1] + J[8] * L[2];
    }
    if (first) {
      // jacobian_scaling = 1 / (1 + sqrt(squared column norm)) from the UNSCALED Jacobian
      // (scale was all ones for this pass); the caller re-linearises with it afterwards
      d.scale[3 * i + 0] = 1.0 / (1.0 + sqrt(h[0]));
      d.scale[3 * i + 1] = 1.0 / (1.0 + sqrt(h[3]));
      d.scale[3 * i + 2] = 1.0 / (1.0 + sqrt(h[5]));
      continue;
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) d.Hd[6 * i + k] = h[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) d.g[3 * i + k] = g[k];
    if (d.is_free[i]) {
      const double * x = d.x + 3 * i, * sc = d.scale + 3 * i;
      // unscaled gradient g / s ; Plus(x, -g): x,y plain, theta wrapped
      const double g0 = g[0] / sc[0], g1 = g[1] / sc[1], g2 = g[2] / sc[2];
      gmax = fmax(gmax, fabs(x[0] - (x[0] - g0)));
      gmax = fmax(gmax, fabs(x[1] - (x[1] - g1)));
      gmax = fmax(gmax, fabs(x[2] - wrap_angle(x[2] - g2)));
      sums[0] += x[0] * x[0] + x[1] * x[1] + x[2] * x[2];
    }
  }
  if (first) return;
  block_sum<1>(sums, red);
  gmax = warp_max(gmax);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[32 + (threadIdx.x >> 5)] = gmax;
  __syncthreads();
  if (threadIdx.x == 0) {
    double m = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) m = fmax(m, red[32 + w]);
    d.partial[blockIdx.x] = sums[0];
    d.partial[kMaxPartials + blockIdx.x] = m;
  }
}

// LM diagonal: clamp(squared column norms) (LevenbergMarquardtStrategy::ComputeStep)
__global__ void k_pg_diag(PgDev d, double min_diag, double max_diag)
{
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < d.N; i += gridDim.x * blockDim.x) {
    d.diag[3 * i + 0] = fmin(fmax(d.Hd[6 * i + 0], min_diag), max_diag);
    d.diag[3 * i + 1] = fmin(fmax(d.Hd[6 * i + 3], min_diag), max_diag);
    d.diag[3 * i + 2] = fmin(fmax(d.Hd[6 * i + 5], min_diag), max_diag);
  }
}

// final reduction of per-CTA partials in fixed order (one warp)
__global__ void k_pg_reduce(PgDev d, int nparts, int slot_sum, int slot_max)
{
  if (threadIdx.x != 0) return;
  double s = 0, m = 0;
  for (int i = 0; i < nparts; ++i) { s += d.partial[i]; m = fmax(m, d.partial[kMaxPartials + i]); }
  if (slot_sum >= 0) d.scalars[slot_sum] = s;
  if (slot_max >= 0) d.scalars[slot_max] = m;
}

// y_i = sum_j A_ij v_j for the scaled normal matrix plus damping: (Hd_i + D_i^2) v_i + sum_e M v_other
__device__ __forceinline__ void spmv_row(const PgDev & d, int i, const double * __restrict__ v0,
                                         const double * __restrict__ v1, double beta, double inv_radius, double out[3],
                                         double vi[3])
{
  // effective vector v = v0 + beta * v1 (v1 may be null)
  auto ld = [&](int j, double w[3]) {
    w[0] = v0[3 * j]; w[1] = v0[3 * j + 1]; w[2] = v0[3 * j + 2];
    if (v1) { w[0] += beta * v1[3 * j]; w[1] += beta * v1[3 * j + 1]; w[2] += beta * v1[3 * j + 2]; }
  };
  ld(i, vi);
  const double * h = d.Hd + 6 * i, * dg = d.diag + 3 * i;
  out[0] = (h[0] + dg[0] * inv_radius) * vi[0] + h[1] * vi[1] + h[2] * vi[2];
  out[1] = h[1] * vi[0] + (h[3] + dg[1] * inv_radius) * vi[1] + h[4] * vi[2];
  out[2] = h[2] * vi[0] + h[4] * vi[1] + (h[5] + dg[2] * inv_radius) * vi[2];
  for (int k = d.adj_start[i]; k < d.adj_start[i + 1]; ++k) {
    const int e = d.adj[k] >> 1, side = d.adj[k] & 1;
    const double * M = d.lin + (size_t)kLin * e + 21;
    const int other = d.eidx[2 * e + (side ? 0 : 1)];
    double w[3];
    ld(other, w);
    if (side == 0) {   // row block a: M w_b
      out[0] += M[0] * w[0] + M[1] * w[1] + M[2] * w[2];
      out[1] += M[3] * w[0] + M[4] * w[1] + M[5] * w[2];
      out[2] += M[6] * w[0] + M[7] * w[1] + M[8] * w[2];
    } else {           // row block b: M^T w_a
      out[0] += M[0] * w[0] + M[3] * w[1] + M[6] * w[2];
      out[1] += M[1] * w[0] + M[4] * w[1] + M[7] * w[2];
      out[2] += M[2] * w[0] + M[5] * w[1] + M[8] * w[2];
    }
  }
}

// Sum of the per-CTA partials of one slot. Warp 0 of every CTA adds them in the same fixed order,
// so every CTA obtains the identical value (deterministic, no atomics); broadcast through smem.
__device__ __forceinline__ double grid_total(const PgDev & d, int slot, int nblk, double * sh)
{
  if (threadIdx.x < 32) {
    const volatile double * p = d.partial + (size_t)slot * kMaxPartials;
    double s = 0;
    for (int i = threadIdx.x; i < nblk; i += 32) s += p[i];
    s = warp_sum(s);
    if (threadIdx.x == 0) sh[0] = s;
  }
  __syncthreads();
  const double r = sh[0];
  __syncthreads();
  return r;
}

// Persistent cooperative PCG: solves (J~^T J~ + D^2/radius) y = J~^T r with the block-Jacobi
// preconditioner M_i = (Hd_i + D_i^2/radius)^-1.  Constant / unused nodes have zero Jacobian
// columns, so their rows reduce to D^2 y = 0.  scalars[8] = iterations, [9] = final relative residual.
__global__ void __launch_bounds__(kPgThreads) k_pg_pcg(PgDev d, double inv_radius, double tol, int max_iter)
{
  cg::grid_group grid = cg::this_grid();
  __shared__ double red[4 * 32];
  __shared__ double bc[1];
  const int nblk = gridDim.x;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;

  // prologue: Minv, r = b, y = 0, z = Minv r, partials of b.b and r.z
  double acc[2] = {0, 0};
  for (int i = tid; i < d.N; i += nth) {
    const double * h = d.Hd + 6 * i, * dg = d.diag + 3 * i;
    const double a00 = h[0] + dg[0] * inv_radius, a01 = h[1], a02 = h[2], a11 = h[3] + dg[1] * inv_radius, a12 = h[4],
                 a22 = h[5] + dg[2] * inv_radius;
    const double c00 = a11 * a22 - a12 * a12, c01 = a02 * a12 - a01 * a22, c02 = a01 * a12 - a02 * a11;
    const double c11 = a00 * a22 - a02 * a02, c12 = a01 * a02 - a00 * a12, c22 = a00 * a11 - a01 * a01;
    const double det = a00 * c00 + a01 * c01 + a02 * c02;
    const double id = 1.0 / det;
    double * mi = d.Minv + 6 * i;
    mi[0] = c00 * id; mi[1] = c01 * id; mi[2] = c02 * id; mi[3] = c11 * id; mi[4] = c12 * id; mi[5] = c22 * id;
    const double b0 = d.g[3 * i], b1 = d.g[3 * i + 1], b2 = d.g[3 * i + 2];
    d.pr[3 * i] = b0; d.pr[3 * i + 1] = b1; d.pr[3 * i + 2] = b2;
    d.y[3 * i] = 0; d.y[3 * i + 1] = 0; d.y[3 * i + 2] = 0;
    const double z0 = mi[0] * b0 + mi[1] * b1 + mi[2] * b2, z1 = mi[1] * b0 + mi[3] * b1 + mi[4] * b2,
                 z2 = mi[2] * b0 + mi[4] * b1 + mi[5] * b2;
    d.pz[3 * i] = z0; d.pz[3 * i + 1] = z1; d.pz[3 * i + 2] = z2;
    d.pp0[3 * i] = 0; d.pp0[3 * i + 1] = 0; d.pp0[3 * i + 2] = 0;
    acc[0] += b0 * b0 + b1 * b1 + b2 * b2;
    acc[1] += b0 * z0 + b1 * z1 + b2 * z2;
  }
  block_sum<2>(acc, red);
  if (threadIdx.x == 0) { d.partial[blockIdx.x] = acc[0]; d.partial[kMaxPartials + blockIdx.x] = acc[1]; }
  grid.sync();
  const double bb = grid_total(d, 0, nblk, bc);
  double rz = grid_total(d, 1, nblk, bc);
  double rr = bb;
  const double stop = tol * tol * bb;
  int it = 0;
  double beta = 0.0;
  double * p_old = d.pp0, * p_new = d.pp1;
  if (bb > 0.0) {
    while (it < max_iter) {
      // phase A: p_new = z + beta p_old ; q = A p_new ; partial p.q
      double a1[1] = {0};
      for (int i = tid; i < d.N; i += nth) {
        double q[3], pi[3];
        spmv_row(d, i, d.pz, p_old, beta, inv_radius, q, pi);
        p_new[3 * i] = pi[0]; p_new[3 * i + 1] = pi[1]; p_new[3 * i + 2] = pi[2];
        d.pq[3 * i] = q[0]; d.pq[3 * i + 1] = q[1]; d.pq[3 * i + 2] = q[2];
        a1[0] += pi[0] * q[0] + pi[1] * q[1] + pi[2] * q[2];
      }
      block_sum<1>(a1, red);
      if (threadIdx.x == 0) d.partial[2 * kMaxPartials + blockIdx.x] = a1[0];
      grid.sync();
      const double pq = grid_total(d, 2, nblk, bc);
      const double alpha = rz / pq;
      // phase B: y += alpha p ; r -= alpha q ; z = Minv r ; partials r.z, r.r
      double a2[2] = {0, 0};
      for (int i = tid; i < d.N; i += nth) {
        double r0 = d.pr[3 * i] - alpha * d.pq[3 * i], r1 = d.pr[3 * i + 1] - alpha * d.pq[3 * i + 1],
               r2 = d.pr[3 * i + 2] - alpha * d.pq[3 * i + 2];
        d.y[3 * i] += alpha * p_new[3 * i]; d.y[3 * i + 1] += alpha * p_new[3 * i + 1]; d.y[3 * i + 2] += alpha * p_new[3 * i + 2];
        d.pr[3 * i] = r0; d.pr[3 * i + 1] = r1; d.pr[3 * i + 2] = r2;
        const double * mi = d.Minv + 6 * i;
        const double z0 = mi[0] * r0 + mi[1] * r1 + mi[2] * r2, z1 = mi[1] * r0 + mi[3] * r1 + mi[4] * r2,
                     z2 = mi[2] * r0 + mi[4] * r1 + mi[5] * r2;
        d.pz[3 * i] = z0; d.pz[3 * i + 1] = z1; d.pz[3 * i + 2] = z2;
        a2[0] += r0 * z0 + r1 * z1 + r2 * z2;
        a2[1] += r0 * r0 + r1 * r1 + r2 * r2;
      }
      block_sum<2>(a2, red);
      // alternate partial slots so a fast CTA cannot overwrite values a slow one still reads
      const int s0 = 3 + 2 * (it & 1);
      if (threadIdx.x == 0) { d.partial[s0 * kMaxPartials + blockIdx.x] = a2[0]; d.partial[(s0 + 1) * kMaxPartials + blockIdx.x] = a2[1]; }
      grid.sync();
      const double rz_new = grid_total(d, s0, nblk, bc);
      rr = grid_total(d, s0 + 1, nblk, bc);
      ++it;
      if (!(rr > stop) || !(pq > 0.0)) break;
      beta = rz_new / rz;
      rz = rz_new;
      double * t = p_old; p_old = p_new; p_new = t;
    }
  }
  if (tid == 0) {
    d.scalars[8] = (double)it;
    d.scalars[9] = bb > 0.0 ? sqrt(rr / bb) : 0.0;
  }
}

// ------------------------------------------------------------------------------------------
// k_pg_pcg_smem: the same PCG, restructured for graphs whose per-CTA share fits shared memory
// (cfg4: 68 nodes / ~550 off-diagonal blocks per CTA).  Each CTA owns a contiguous range of nodes and
// keeps THEIR rows of the block-sparse normal matrix (3x3 blocks, already oriented), the diagonal
// blocks, the preconditioner and all CG vectors of its nodes in shared memory for the whole solve.
// Per CG iteration the only global traffic is the neighbour gather of z and p (L2) and 2 light
// grid barriers (one atomic counter; partial dot products are summed by every CTA in the same fixed
// order, so the result is bit-reproducible).
// ------------------------------------------------------------------------------------------
struct PcgSmemCfg {
  int npc;          // nodes per CTA
  int max_slots;    // max off-diagonal blocks of one CTA
  double * gz;      // [N][3]
  double * gp;      // [2][N][3]
  unsigned int * bar;   // barrier counter (zeroed before launch)
};

__device__ __forceinline__ void grid_barrier(unsigned int * bar, unsigned int target)
{
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar, 1u);
    unsigned int v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
    } while (v < target);
  }
  __syncthreads();
}

__device__ __forceinline__ double ld_cg(const double * p) { return __ldcg(p); }

__global__ void __launch_bounds__(512, 1) k_pg_pcg_smem(PgDev d, PcgSmemCfg c, double inv_radius, double tol, int max_iter)
{
  extern __shared__ __align__(16) unsigned char sm_raw[];
  __shared__ double red[4 * 32];
  __shared__ double bc[1];
  const int T = blockDim.x, tid = threadIdx.x, G = gridDim.x;
  const int lo = min(d.N, (int)blockIdx.x * c.npc), hi = min(d.N, lo + c.npc), nloc = hi - lo;
  const int s_lo = d.adj_start[lo], nslots = d.adj_start[hi] - s_lo;
  double * sB = reinterpret_cast<double *>(sm_raw);            // [max_slots][9]
  double * sV = sB + (size_t)c.max_slots * 9;                  // [max_slots][3]
  double * sH = sV + (size_t)c.max_slots * 3;                  // [npc][6]
  double * sMi = sH + (size_t)c.npc * 6;                       // [npc][6]
  double * sR = sMi + (size_t)c.npc * 6;                       // [npc][3] each below
  double * sZ = sR + (size_t)c.npc * 3;
  double * sP = sZ + (size_t)c.npc * 3;
  double * sQ = sP + (size_t)c.npc * 3;
  double * sY = sQ + (size_t)c.npc * 3;
  int * sCol = reinterpret_cast<int *>(sY + (size_t)c.npc * 3);   // [max_slots]
  int * sStart = sCol + c.max_slots;                              // [npc + 1]
  unsigned int bar_target = 0;

  // ---- prologue: load this CTA's rows ----
  for (int i = tid; i <= nloc; i += T) sStart[i] = d.adj_start[lo + i] - s_lo;
  for (int s = tid; s < nslots; s += T) {
    const int a = d.adj[s_lo + s];
    const int e = a >> 1, side = a & 1;
    const double * M = d.lin + (size_t)kLin * e + 21;
    sCol[s] = d.eidx[2 * e + (side ? 0 : 1)];
    double * B = sB + 9 * s;
    if (side == 0) {
#pragma unroll
      for (int k = 0; k < 9; ++k) B[k] = M[k];
    } else {
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) B[3 * i + j] = M[3 * j + i];
    }
  }
  double acc[2] = {0, 0};
  for (int n = tid; n < nloc; n += T) {
    const int i = lo + n;
    const double * h = d.Hd + 6 * i, * dg = d.diag + 3 * i;
    const double a00 = h[0] + dg[0] * inv_radius, a01 = h[1], a02 = h[2], a11 = h[3] + dg[1] * inv_radius, a12 = h[4],
                 a22 = h[5] + dg[2] * inv_radius;
    double * H = sH + 6 * n;
    H[0] = a00; H[1] = a01; H[2] = a02; H[3] = a11; H[4] = a12; H[5] = a22;
    const double c00 = a11 * a22 - a12 * a12, c01 = a02 * a12 - a01 * a22, c02 = a01 * a12 - a02 * a11;
    const double c11 = a00 * a22 - a02 * a02, c12 = a01 * a02 - a00 * a12, c22 = a00 * a11 - a01 * a01;
    const double id = 1.0 / (a00 * c00 + a01 * c01 + a02 * c02);
    double * mi = sMi + 6 * n;
    mi[0] = c00 * id; mi[1] = c01 * id; mi[2] = c02 * id; mi[3] = c11 * id; mi[4] = c12 * id; mi[5] = c22 * id;
    const double b0 = d.g[3 * i], b1 = d.g[3 * i + 1], b2 = d.g[3 * i + 2];
    const double z0 = mi[0] * b0 + mi[1] * b1 + mi[2] * b2, z1 = mi[1] * b0 + mi[3] * b1 + mi[4] * b2,
                 z2 = mi[2] * b0 + mi[4] * b1 + mi[5] * b2;
    sR[3 * n] = b0; sR[3 * n + 1] = b1; sR[3 * n + 2] = b2;
    sZ[3 * n] = z0; sZ[3 * n + 1] = z1; sZ[3 * n + 2] = z2;
    sP[3 * n] = 0; sP[3 * n + 1] = 0; sP[3 * n + 2] = 0;
    sY[3 * n] = 0; sY[3 * n + 1] = 0; sY[3 * n + 2] = 0;
    c.gz[3 * i] = z0; c.gz[3 * i + 1] = z1; c.gz[3 * i + 2] = z2;
    c.gp[3 * i] = 0; c.gp[3 * i + 1] = 0; c.gp[3 * i + 2] = 0;
    acc[0] += b0 * b0 + b1 * b1 + b2 * b2;
    acc[1] += b0 * z0 + b1 * z1 + b2 * z2;
  }
  block_sum<2>(acc, red);
  if (tid == 0) { d.partial[blockIdx.x] = acc[0]; d.partial[kMaxPartials + blockIdx.x] = acc[1]; }
  bar_target += G;
  grid_barrier(c.bar, bar_target);
  const double bb = grid_total(d, 0, G, bc);
  double rz = grid_total(d, 1, G, bc);
  double rr = bb;
  const double stop = tol * tol * bb;
  int it = 0;
  double beta = 0.0;
  int cur = 0;   // gp[cur] holds p_old
  if (bb > 0.0) {
    while (it < max_iter) {
      const double * gpo = c.gp + (size_t)cur * 3 * d.N;
      double * gpn = c.gp + (size_t)(cur ^ 1) * 3 * d.N;
      // ---- phase A: gather v_j = z_j + beta p_j ; p_new ; q = A p_new ; partial p.q ----
      for (int s = tid; s < nslots; s += T) {
        const int j = sCol[s];
        double v0, v1, v2;
        if (j >= lo && j < hi) {   // neighbour owned by this CTA: shared memory
          const int n = j - lo;
          v0 = sZ[3 * n] + beta * sP[3 * n]; v1 = sZ[3 * n + 1] + beta * sP[3 * n + 1]; v2 = sZ[3 * n + 2] + beta * sP[3 * n + 2];
        } else {
          v0 = ld_cg(c.gz + 3 * j) + beta * ld_cg(gpo + 3 * j);
          v1 = ld_cg(c.gz + 3 * j + 1) + beta * ld_cg(gpo + 3 * j + 1);
          v2 = ld_cg(c.gz + 3 * j + 2) + beta * ld_cg(gpo + 3 * j + 2);
        }
        sV[3 * s] = v0; sV[3 * s + 1] = v1; sV[3 * s + 2] = v2;
      }
      __syncthreads();
      for (int k = tid; k < 3 * nloc; k += T) {   // new search direction of own nodes (after the local reads above)
        const double pn = sZ[k] + beta * sP[k];
        sQ[k] = pn;   // temporarily: p_new
      }
      __syncthreads();
      for (int k = tid; k < 3 * nloc; k += T) { sP[k] = sQ[k]; gpn[3 * lo + k] = sQ[k]; }
      __syncthreads();
      double a1[1] = {0};
      for (int k = tid; k < 3 * nloc; k += T) {
        const int n = k / 3, r = k - 3 * n;
        const double * H = sH + 6 * n;
        const double p0 = sP[3 * n], p1 = sP[3 * n + 1], p2 = sP[3 * n + 2];
        double q;
        if (r == 0) q = H[0] * p0 + H[1] * p1 + H[2] * p2;
        else if (r == 1) q = H[1] * p0 + H[3] * p1 + H[4] * p2;
        else q = H[2] * p0 + H[4] * p1 + H[5] * p2;
        for (int s = sStart[n]; s < sStart[n + 1]; ++s) {
          const double * B = sB + 9 * s + 3 * r, * v = sV + 3 * s;
          q += B[0] * v[0] + B[1] * v[1] + B[2] * v[2];
        }
        sQ[k] = q;
        a1[0] += sP[k] * q;
      }
      block_sum<1>(a1, red);
      if (tid == 0) d.partial[2 * kMaxPartials + blockIdx.x] = a1[0];
      bar_target += G;
      grid_barrier(c.bar, bar_target);
      const double pq = grid_total(d, 2, G, bc);
      const double alpha = rz / pq;
      // ---- phase B: y += alpha p ; r -= alpha q ; z = Minv r ; partials r.z, r.r ----
      for (int k = tid; k < 3 * nloc; k += T) { sY[k] += alpha * sP[k]; sR[k] -= alpha * sQ[k]; }
      __syncthreads();
      double a2[2] = {0, 0};
      for (int k = tid; k < 3 * nloc; k += T) {
        const int n = k / 3, r = k - 3 * n;
        const double * mi = sMi + 6 * n;
        const double r0 = sR[3 * n], r1 = sR[3 * n + 1], r2 = sR[3 * n + 2];
        double z;
        if (r == 0) z = mi[0] * r0 + mi[1] * r1 + mi[2] * r2;
        else if (r == 1) z = mi[1] * r0 + mi[3] * r1 + mi[4] * r2;
        else z = mi[2] * r0 + mi[4] * r1 + mi[5] * r2;
        sZ[k] = z;
        c.gz[3 * lo + k] = z;
        a2[0] += sR[k] * z;
        a2[1] += sR[k] * sR[k];
      }
      block_sum<2>(a2, red);
      const int s0 = 3 + 2 * (it & 1);
      if (tid == 0) { d.partial[s0 * kMaxPartials + blockIdx.x] = a2[0]; d.partial[(s0 + 1) * kMaxPartials + blockIdx.x] = a2[1]; }
      bar_target += G;
      grid_barrier(c.bar, bar_target);
      const double rz_new = grid_total(d, s0, G, bc);
      rr = grid_total(d, s0 + 1, G, bc);
      ++it;
      cur ^= 1;
      if (!(rr > stop) || !(pq > 0.0)) break;
      beta = rz_new / rz;
      rz = rz_new;
    }
  }
  for (int k = tid; k < 3 * nloc; k += T) d.y[3 * lo + k] = sY[k];
  if (blockIdx.x == 0 && tid == 0) {
    d.scalars[8] = (double)it;
    d.scalars[9] = bb > 0.0 ? sqrt(rr / bb) : 0.0;
  }
}

// ------------------------------------------------------------------------------------------
// k_pg_pcg_2lvl: shared-memory-resident PCG (as k_pg_pcg_smem) with a TWO-LEVEL preconditioner
//     M^-1 = blockdiag(H_ii + D_i)^-1  +  P Ac^-1 P^T ,   Ac = P^T (H + D) P
// One aggregate per CTA (its contiguous node range); P holds CM coarse modes per aggregate, expressed
// in the Jacobi-scaled variables:
//   CM = 3  the aggregate's rigid-body modes (translation x, y, rotation about its centroid)
//   CM = 6  the same three modes once more, weighted by s in [-1, 1] = the node's position along the
//           aggregate (aggregates are stretches of the trajectory): the piecewise-LINEAR deformation
//           modes.  At cfg4 this costs 1.6x fewer CG iterations than CM = 3 (tools/precond_study.py:
//           359 vs 568 at LM step 2) for a coarse matrix of 888 instead of 444 rows.
// The low-frequency deformation modes that make block-Jacobi CG need thousands of iterations on a
// pose graph are removed by the coarse solve.
//   setup per solve: every CTA builds its CM rows of Ac, then a block Gauss-Jordan over the grid
//     (one CM-row pivot exchange per aggregate) leaves each CTA holding ITS CM rows of Ac^-1 in smem;
//   per CG iteration: 2 flag-based exchanges (no atomic barrier): {p.q, P^T q} and {r.z, r.r};
//     the coarse residual P^T r is carried by the recurrence rc -= alpha P^T q, identically in
//     every CTA, so the coarse correction needs no extra exchange.
// All reductions are summed in a fixed order: results are bit-reproducible.
// ------------------------------------------------------------------------------------------
struct Pcg2Cfg {
  int npc, max_slots;     // npc = MAX nodes of one aggregate (array sizing)
  int ex_doubles;         // size of the exchange scratch (>= (1 + CM) G and large enough for the set-up alias)
  const int32_t * agg_start;   // [G + 1] contiguous node ranges of equal node count
  const int32_t * agg_of;      // [N] aggregate of every node
  double * gz;            // [N][3]
  double * gp;            // [2][N][3]
  double * gPt;           // [N][10]  P~ base block of every node (rows: node comps, cols: rigid modes) + its s
  double * gRow;          // [G][CM][2 nc] published pivot rows of the block Gauss-Jordan
  double * grc;           // [G][CM]  initial coarse residual
  double * e1;            // [2][G] slots: {p.q partial, P^T q (CM)}
  double * e2;            // [2][G] slots: {r.z partial, r.r partial}
  unsigned int * gjflag;  // [G]
  unsigned int * bar;     // atomic barrier counter (set-up only)
};

__device__ __forceinline__ double pg_sentinel() { return __longlong_as_double(0x7FF8DEADBEEF0001LL); }
__device__ __forceinline__ bool pg_is_sentinel(double v) { return __double_as_longlong(v) == 0x7FF8DEADBEEF0001LL; }

// Exchange slots: one 256-byte line per (parity, CTA) so that the all-to-all polling spreads over
// every L2 slice instead of hammering a few sectors; a slot holds K <= 8 doubles, each self-flagged
// (a value is "published" when it is not the sentinel).  One thread per source CTA polls with
// 16-byte loads.
constexpr int kSlotStride = 32;   // doubles
__device__ __forceinline__ void ld_volatile2(const double * p, double & a, double & b)
{
  asm volatile("ld.volatile.global.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "l"(p) : "memory");
}
template <int K>
__device__ __forceinline__ void poll_slots(const double * slots, int G, double * s_out)
{
  constexpr int K2 = (K + 1) / 2;
  for (int t = threadIdx.x; t < G; t += blockDim.x) {
    const double * p = slots + (size_t)t * kSlotStride;
    double v[2 * K2];
    bool ok;
    do {
#pragma unroll
      for (int k = 0; k < K2; ++k) ld_volatile2(p + 2 * k, v[2 * k], v[2 * k + 1]);
      ok = true;
#pragma unroll
      for (int k = 0; k < K; ++k) ok = ok && !pg_is_sentinel(v[k]);
    } while (!ok);
#pragma unroll
    for (int k = 0; k < K; ++k) s_out[t * K + k] = v[k];
    __threadfence();
  }
  __syncthreads();
}
// fixed-order sum of s[i * stride + off], i < n, by warp 0; broadcast through bc
__device__ __forceinline__ double ordered_sum(const double * s, int n, int stride, int off, double * bc)
{
  if (threadIdx.x < 32) {
    double a = 0;
    for (int i = threadIdx.x; i < n; i += 32) a += s[i * stride + off];
    a = warp_sum(a);
    if (threadIdx.x == 0) bc[0] = a;
  }
  __syncthreads();
  const double r = bc[0];
  __syncthreads();
  return r;
}

template <int CM>
__global__ void __launch_bounds__(256, 2) k_pg_pcg_2lvl(PgDev d, Pcg2Cfg c, double inv_radius, double tol, int max_iter)
{
  static_assert(CM == 3 || CM == 6, "coarse modes per aggregate");
  constexpr int KE1 = 1 + CM;            // doubles of an E1 slot
  extern __shared__ __align__(16) unsigned char sm_raw[];
  __shared__ double red[KE1 * 32];
  __shared__ double bc[1];
  __shared__ double s_small[CM * CM + 4];
  __shared__ double s_y[8];
  const int T = blockDim.x, tid = threadIdx.x, G = gridDim.x, I = blockIdx.x;
  const int nc = CM * G;
  const int lo = c.agg_start[I], hi = c.agg_start[I + 1], nloc = hi - lo;
  const int s_lo = d.adj_start[lo], nslots = d.adj_start[hi] - s_lo;
  double * sB = reinterpret_cast<double *>(sm_raw);            // [max_slots][9]
  double * sV = sB + (size_t)c.max_slots * 9;                  // [max_slots][3]
  double * sEx = sV + (size_t)c.max_slots * 3;                 // [ex_doubles >= KE1 G] exchange scratch, right after sV
  double * sH = sEx + (size_t)c.ex_doubles;                    // [npc][6]
  double * sMi = sH + (size_t)c.npc * 6;                       // [npc][6]
  double * sR = sMi + (size_t)c.npc * 6;                       // [npc][3] each below
  double * sZ = sR + (size_t)c.npc * 3;
  double * sP = sZ + (size_t)c.npc * 3;
  double * sQ = sP + (size_t)c.npc * 3;
  double * sY = sQ + (size_t)c.npc * 3;
  double * sPt = sY + (size_t)c.npc * 3;                       // [npc][9]
  double * sS = sPt + (size_t)c.npc * 9;                       // [npc] position of the node along its aggregate, [-1, 1]
  double * sAr = sS + (size_t)c.npc;                           // [CM][nc] right half of [Ac | I] -> rows of Ac^-1
  double * sRc = sAr + (size_t)CM * nc;                        // [nc] coarse residual (identical in all CTAs)
  // the left half of [Ac | I] only lives during set-up: it aliases the CG-only scratch sV | sEx
  // (3 max_slots + ex_doubles >= CM nc is guaranteed by the host)
  double * sAl = sV;                                           // [CM][nc]
  int * sCol = reinterpret_cast<int *>(sRc + nc);              // [max_slots]
  int * sNode = sCol + c.max_slots;                            // [max_slots] local node of the slot
  int * sStart = sNode + c.max_slots;                          // [npc + 1]
  unsigned int bar_target = 0;
  const double SENT = pg_sentinel();
  unsigned long long t_start = 0, t_setup = 0, t_gj = 0;
  unsigned long long tA = 0, tB = 0, tC = 0, tD = 0, tE = 0, t0 = 0, t1 = 0;   // CG phase timers
#define PG_TICK(acc) do { if (I == 0 && tid == 0) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); acc += t1 - t0; t0 = t1; } } while (0)
  if (I == 0 && tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));

  // ---- load this CTA's rows (as k_pg_pcg_smem) ----
  for (int i = tid; i <= nloc; i += T) sStart[i] = d.adj_start[lo + i] - s_lo;
  __syncthreads();
  for (int n = tid; n < nloc; n += T)
    for (int s = sStart[n]; s < sStart[n + 1]; ++s) sNode[s] = n;
  for (int s = tid; s < nslots; s += T) {
    const int a = d.adj[s_lo + s];
    const int e = a >> 1, side = a & 1;
    const double * M = d.lin + (size_t)kLin * e + 21;
    sCol[s] = d.eidx[2 * e + (side ? 0 : 1)];
    double * B = sB + 9 * s;
    if (side == 0) {
#pragma unroll
      for (int k = 0; k < 9; ++k) B[k] = M[k];
    } else {
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) B[3 * i + j] = M[3 * j + i];
    }
  }
  // centroid of the aggregate's free nodes (fixed-order sum by thread 0: nloc is small)
  if (tid == 0) {
    double cx = 0, cy = 0; int cnt = 0;
    for (int n = 0; n < nloc; ++n)
      if (d.is_free[lo + n]) { cx += d.x[3 * (lo + n)]; cy += d.x[3 * (lo + n) + 1]; ++cnt; }
    s_small[0] = cnt ? cx / cnt : 0.0; s_small[1] = cnt ? cy / cnt : 0.0;
    s_small[2] = (double)cnt;
  }
  __syncthreads();
  for (int n = tid; n < nloc; n += T) {
    const int i = lo + n;
    const double * h = d.Hd + 6 * i, * dg = d.diag + 3 * i;
    const double a00 = h[0] + dg[0] * inv_radius, a01 = h[1], a02 = h[2], a11 = h[3] + dg[1] * inv_radius, a12 = h[4],
                 a22 = h[5] + dg[2] * inv_radius;
    double * H = sH + 6 * n;
    H[0] = a00; H[1] = a01; H[2] = a02; H[3] = a11; H[4] = a12; H[5] = a22;
    const double c00 = a11 * a22 - a12 * a12, c01 = a02 * a12 - a01 * a22, c02 = a01 * a12 - a02 * a11;
    const double c11 = a00 * a22 - a02 * a02, c12 = a01 * a02 - a00 * a12, c22 = a00 * a11 - a01 * a01;
    const double id = 1.0 / (a00 * c00 + a01 * c01 + a02 * c02);
    double * mi = sMi + 6 * n;
    mi[0] = c00 * id; mi[1] = c01 * id; mi[2] = c02 * id; mi[3] = c11 * id; mi[4] = c12 * id; mi[5] = c22 * id;
    // P~ base block: rigid-body modes of the aggregate in Jacobi-scaled variables (y~ = y / s); zero for constant nodes
    double * pt = sPt + 9 * n;
    const double f = d.is_free[i] ? 1.0 : 0.0;
    const double isx = f / d.scale[3 * i], isy = f / d.scale[3 * i + 1], ist = f / d.scale[3 * i + 2];
    pt[0] = isx; pt[1] = 0;   pt[2] = -(d.x[3 * i + 1] - s_small[1]) * isx;
    pt[3] = 0;   pt[4] = isy; pt[5] = (d.x[3 * i] - s_small[0]) * isy;
    pt[6] = 0;   pt[7] = 0;   pt[8] = ist;
    // the s-weighted modes need two free nodes to be independent of the rigid ones; otherwise they are switched off
    // (s = 0 gives zero rows and columns of Ac, which the Gauss-Jordan replaces by the identity)
    const double sn = (CM > 3 && nloc > 1 && s_small[2] >= 2.0) ? 2.0 * n / (double)(nloc - 1) - 1.0 : 0.0;
    sS[n] = sn;
#pragma unroll
    for (int k = 0; k < 9; ++k) c.gPt[10 * (size_t)i + k] = pt[k];
    c.gPt[10 * (size_t)i + 9] = sn;
  }
  // own exchange slots start empty
  if (tid < 2 * KE1) c.e1[((size_t)(tid / KE1) * G + I) * kSlotStride + (tid % KE1)] = SENT;
  if (tid < 4) c.e2[((size_t)(tid >> 1) * G + I) * kSlotStride + (tid & 1)] = SENT;
  bar_target += G;
  grid_barrier(c.bar, bar_target);   // gPt, empty slots visible everywhere

  // ---- coarse operator: this CTA's CM rows of Ac = P^T (H + D) P, then block Gauss-Jordan ----
  // With P_i = [pi | s_i pi], the (I, ct) block of Ac is [[W, Wj], [Wi, Wij]] with W = sum pi^T A_ij pj and the
  // sums weighted by s_j, s_i, s_i s_j.
  for (int k = tid; k < CM * nc; k += T) { sAl[k] = 0.0; sAr[k] = 0.0; }
  __syncthreads();
  for (int ct = tid; ct < G; ct += T) {   // a thread owns coarse column block ct; slots are visited in order: deterministic
    double acc[CM == 3 ? 9 : 36];
#pragma unroll
    for (int k = 0; k < (CM == 3 ? 9 : 36); ++k) acc[k] = 0.0;
    auto add_block = [&](const double (&w3)[9], double si, double sj) {   // w3 = pi^T A pj
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        acc[k] += w3[k];
        if constexpr (CM > 3) { acc[9 + k] += sj * w3[k]; acc[18 + k] += si * w3[k]; acc[27 + k] += si * sj * w3[k]; }
      }
    };
    for (int s = 0; s < nslots; ++s) {
      const int j = sCol[s];
      if (c.agg_of[j] != ct) continue;
      const double * B = sB + 9 * s, * pi = sPt + 9 * sNode[s];
      double pj[9], w[9], w3[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) pj[k] = ld_cg(c.gPt + 10 * (size_t)j + k);
      const double sj = ld_cg(c.gPt + 10 * (size_t)j + 9);
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int q = 0; q < 3; ++q) w[3 * r + q] = B[3 * r] * pj[q] + B[3 * r + 1] * pj[3 + q] + B[3 * r + 2] * pj[6 + q];
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int q = 0; q < 3; ++q) w3[3 * r + q] = pi[r] * w[q] + pi[3 + r] * w[3 + q] + pi[6 + r] * w[6 + q];
      add_block(w3, sS[sNode[s]], sj);
    }
    if (ct == I) {   // diagonal blocks of own nodes
      for (int n = 0; n < nloc; ++n) {
        const double * H = sH + 6 * n, * pi = sPt + 9 * n;
        double w[9], w3[9];
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          w[q] = H[0] * pi[q] + H[1] * pi[3 + q] + H[2] * pi[6 + q];
          w[3 + q] = H[1] * pi[q] + H[3] * pi[3 + q] + H[4] * pi[6 + q];
          w[6 + q] = H[2] * pi[q] + H[4] * pi[3 + q] + H[5] * pi[6 + q];
        }
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int q = 0; q < 3; ++q) w3[3 * r + q] = pi[r] * w[q] + pi[3 + r] * w[3 + q] + pi[6 + r] * w[6 + q];
        add_block(w3, sS[n], sS[n]);
      }
    }
    // acc layout: [rb][cb][r][q] with rb / cb = 0 rigid, 1 s-weighted; Ac row = 3 rb + r, column = CM ct + 3 cb + q
#pragma unroll
    for (int rb = 0; rb < CM / 3; ++rb)
#pragma unroll
      for (int cb = 0; cb < CM / 3; ++cb)
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int q = 0; q < 3; ++q)
            sAl[(size_t)(3 * rb + r) * nc + CM * ct + 3 * cb + q] = acc[(CM == 3 ? 0 : 18 * rb + 9 * cb) + 3 * r + q];
  }
  if (tid < CM) sAr[(size_t)tid * nc + CM * I + tid] = 1.0;   // augmented identity
  __syncthreads();
  if (I == 0 && tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_setup));
  for (int k = 0; k < G; ++k) {
    double * row = c.gRow + (size_t)k * CM * 2 * nc;   // published pivot rows: [CM][2 nc] (left | right)
    auto elem = [&](int r, int j) -> double & { return j < nc ? sAl[(size_t)r * nc + j] : sAr[(size_t)r * nc + (j - nc)]; };
    // columns that can be non-zero in pivot rows k: the not yet reduced part of the left half, [CM k, nc), and the part of
    // the right half filled so far, [0, CM (k + 1)); everything else is 0 and stays 0
    const int nleft = nc - CM * k, nact = nleft + CM * (k + 1);
    if (I == k) {
      // inverse of the CM x CM pivot block by Gauss-Jordan without pivoting (the block is symmetric positive definite on
      // its non-degenerate modes); a mode with a zero diagonal (no free node, or the s-modes of a one-node aggregate) has
      // a zero row and column in all of Ac: it is replaced by the identity
      if (tid == 0) {
        double a[CM][2 * CM];
#pragma unroll
        for (int r = 0; r < CM; ++r)
#pragma unroll
          for (int q = 0; q < CM; ++q) { a[r][q] = elem(r, CM * k + q); a[r][CM + q] = (r == q) ? 1.0 : 0.0; }
        double d0[CM];
#pragma unroll
        for (int p = 0; p < CM; ++p) d0[p] = fabs(a[p][p]);
#pragma unroll
        for (int p = 0; p < CM; ++p) {
          if (!(fabs(a[p][p]) > 1e-12 * d0[p]) || !(d0[p] > 1e-300)) {   // zero or (numerically) dependent mode
#pragma unroll
            for (int q = 0; q < 2 * CM; ++q) a[p][q] = 0.0;
#pragma unroll
            for (int r = 0; r < CM; ++r) a[r][p] = 0.0;
            a[p][p] = 1.0; a[p][CM + p] = 1.0;
          }
          const double ip = 1.0 / a[p][p];
#pragma unroll
          for (int q = 0; q < 2 * CM; ++q) a[p][q] *= ip;
#pragma unroll
          for (int r = 0; r < CM; ++r) {
            if (r == p) continue;
            const double m = a[r][p];
#pragma unroll
            for (int q = 0; q < 2 * CM; ++q) a[r][q] -= m * a[p][q];
          }
        }
#pragma unroll
        for (int r = 0; r < CM; ++r)
#pragma unroll
          for (int q = 0; q < CM; ++q) s_small[CM * r + q] = a[r][CM + q];
      }
      __syncthreads();
      for (int t = tid; t < nact; t += T) {
        const int j = t < nleft ? CM * k + t : nc + (t - nleft);
        double v[CM], o[CM];
#pragma unroll
        for (int r = 0; r < CM; ++r) v[r] = elem(r, j);
#pragma unroll
        for (int r = 0; r < CM; ++r) {
          double a = 0;
#pragma unroll
          for (int q = 0; q < CM; ++q) a += s_small[CM * r + q] * v[q];
          o[r] = a;
        }
#pragma unroll
        for (int r = 0; r < CM; ++r) { elem(r, j) = o[r]; row[(size_t)r * 2 * nc + j] = o[r]; }
      }
      __syncthreads();
      if (tid == 0) {
        __threadfence();
        asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(c.gjflag + k), "r"(1u) : "memory");
      }
    } else {
      if (tid == 0) {
        unsigned int v;
        do {
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(c.gjflag + k) : "memory");
        } while (v == 0u);
      }
      if (tid < CM * CM) s_small[tid] = elem(tid / CM, CM * k + (tid % CM));   // my multipliers (read before they are eliminated)
      __syncthreads();
      for (int t = tid; t < nact; t += T) {
        const int j = t < nleft ? CM * k + t : nc + (t - nleft);
        double pr[CM];
#pragma unroll
        for (int q = 0; q < CM; ++q) pr[q] = ld_cg(row + (size_t)q * 2 * nc + j);
#pragma unroll
        for (int r = 0; r < CM; ++r) {
          double a = 0;
#pragma unroll
          for (int q = 0; q < CM; ++q) a += s_small[CM * r + q] * pr[q];
          elem(r, j) -= a;
        }
      }
      __syncthreads();
    }
  }
  if (I == 0 && tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_gj));
  // rows of Ac^-1 are now sAr[r * nc + j]
  __syncthreads();

  // ---- CG start: r = b, coarse residual, z = M^-1 r ----
  double accb[1] = {0};
  for (int k = tid; k < 3 * nloc; k += T) {
    const double b = d.g[3 * lo + k];
    sR[k] = b; sY[k] = 0.0; sP[k] = 0.0;
    c.gp[3 * lo + k] = 0.0;
    accb[0] += b * b;
  }
  __syncthreads();
  if (tid < CM) {   // P^T r of this aggregate, fixed order
    const int cb = tid / 3, q = tid % 3;
    double a = 0;
    for (int n = 0; n < nloc; ++n) {
      const double v = sPt[9 * n + q] * sR[3 * n] + sPt[9 * n + 3 + q] * sR[3 * n + 1] + sPt[9 * n + 6 + q] * sR[3 * n + 2];
      a += cb ? sS[n] * v : v;
    }
    c.grc[CM * I + tid] = a;
  }
  block_sum<1>(accb, red);
  if (tid == 0) d.partial[I] = accb[0];
  bar_target += G;
  grid_barrier(c.bar, bar_target);
  const double bb = grid_total(d, 0, G, bc);
  for (int k = tid; k < nc; k += T) sRc[k] = ld_cg(c.grc + k);
  __syncthreads();

  // z = blockJacobi^-1 r + P Ac^-1 rc ; returns partial r.z and r.r through a2
  auto apply_precond = [&](double (&a2)[2]) {
    if (tid < 32 * CM) {   // CM warps: one row of Ac^-1 each
      const int w = tid >> 5, l = tid & 31;
      const double * Ai = sAr + (size_t)w * nc;
      double a = 0;
      for (int j = l; j < nc; j += 32) a += Ai[j] * sRc[j];
      a = warp_sum(a);
      if (l == 0) s_y[w] = a;
    }
    __syncthreads();
    a2[0] = 0; a2[1] = 0;
    for (int k = tid; k < 3 * nloc; k += T) {
      const int n = k / 3, r = k - 3 * n;
      const double * mi = sMi + 6 * n, * pt = sPt + 9 * n + 3 * r;
      const double r0 = sR[3 * n], r1 = sR[3 * n + 1], r2 = sR[3 * n + 2];
      double z;
      if (r == 0) z = mi[0] * r0 + mi[1] * r1 + mi[2] * r2;
      else if (r == 1) z = mi[1] * r0 + mi[3] * r1 + mi[4] * r2;
      else z = mi[2] * r0 + mi[4] * r1 + mi[5] * r2;
      double y0 = s_y[0], y1 = s_y[1], y2 = s_y[2];
      if constexpr (CM > 3) { const double sn = sS[n]; y0 += sn * s_y[3]; y1 += sn * s_y[4]; y2 += sn * s_y[5]; }
      z += pt[0] * y0 + pt[1] * y1 + pt[2] * y2;
      sZ[k] = z;
      c.gz[3 * lo + k] = z;
      a2[0] += sR[k] * z;
      a2[1] += sR[k] * sR[k];
    }
  };
  double a2[2];
  apply_precond(a2);
  block_sum<2>(a2, red);
  if (tid == 0) d.partial[kMaxPartials + I] = a2[0];
  bar_target += G;
  grid_barrier(c.bar, bar_target);
  double rz = grid_total(d, 1, G, bc);
  double rr = bb;
  const double stop = tol * tol * bb;
  int it = 0;
  double beta = 0.0;
  int cur = 0;
  if (bb > 0.0) {
    while (it < max_iter) {
      const int par = it & 1;
      const double * gpo = c.gp + (size_t)cur * 3 * d.N;
      double * gpn = c.gp + (size_t)(cur ^ 1) * 3 * d.N;
      double * e1 = c.e1 + (size_t)par * G * kSlotStride, * e2 = c.e2 + (size_t)par * G * kSlotStride;
      // ---- phase A ----
      if (I == 0 && tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
      for (int s = tid; s < nslots; s += T) {
        const int j = sCol[s];
        double v0, v1, v2;
        if (j >= lo && j < hi) {
          const int n = j - lo;
          v0 = sZ[3 * n] + beta * sP[3 * n]; v1 = sZ[3 * n + 1] + beta * sP[3 * n + 1]; v2 = sZ[3 * n + 2] + beta * sP[3 * n + 2];
        } else {
          v0 = ld_cg(c.gz + 3 * j) + beta * ld_cg(gpo + 3 * j);
          v1 = ld_cg(c.gz + 3 * j + 1) + beta * ld_cg(gpo + 3 * j + 1);
          v2 = ld_cg(c.gz + 3 * j + 2) + beta * ld_cg(gpo + 3 * j + 2);
        }
        sV[3 * s] = v0; sV[3 * s + 1] = v1; sV[3 * s + 2] = v2;
      }
      __syncthreads();
      PG_TICK(tA);
      for (int k = tid; k < 3 * nloc; k += T) sQ[k] = sZ[k] + beta * sP[k];
      __syncthreads();
      for (int k = tid; k < 3 * nloc; k += T) { sP[k] = sQ[k]; gpn[3 * lo + k] = sQ[k]; }
      __syncthreads();
      double a1[KE1];   // p.q and the CM components of P^T q of this aggregate
#pragma unroll
      for (int k = 0; k < KE1; ++k) a1[k] = 0.0;
      for (int k = tid; k < 3 * nloc; k += T) {
        const int n = k / 3, r = k - 3 * n;
        const double * H = sH + 6 * n;
        const double p0 = sP[3 * n], p1 = sP[3 * n + 1], p2 = sP[3 * n + 2];
        double q;
        if (r == 0) q = H[0] * p0 + H[1] * p1 + H[2] * p2;
        else if (r == 1) q = H[1] * p0 + H[3] * p1 + H[4] * p2;
        else q = H[2] * p0 + H[4] * p1 + H[5] * p2;
        for (int s = sStart[n]; s < sStart[n + 1]; ++s) {
          const double * B = sB + 9 * s + 3 * r, * v = sV + 3 * s;
          q += B[0] * v[0] + B[1] * v[1] + B[2] * v[2];
        }
        sQ[k] = q;
        const double * pt = sPt + 9 * n + 3 * r;
        a1[0] += sP[k] * q;
        const double u0 = pt[0] * q, u1 = pt[1] * q, u2 = pt[2] * q;
        a1[1] += u0; a1[2] += u1; a1[3] += u2;
        if constexpr (CM > 3) { const double sn = sS[n]; a1[4] += sn * u0; a1[5] += sn * u1; a1[6] += sn * u2; }
      }
      block_sum<KE1>(a1, red);
      PG_TICK(tB);
      if (tid == 0) {
        __threadfence();   // p_new of this CTA visible before the flagged values
        double * m = e1 + (size_t)I * kSlotStride;
#pragma unroll
        for (int k = 1; k < KE1; ++k) m[k] = a1[k];
        m[0] = a1[0];
      }
      poll_slots<KE1>(e1, G, sEx);
      PG_TICK(tC);
      // every CTA published E1(it) only after it finished reading E2(it-1): those slots can be recycled now
      if (it > 0 && tid < 2) c.e2[((size_t)(par ^ 1) * G + I) * kSlotStride + tid] = SENT;
      const double pq = ordered_sum(sEx, G, KE1, 0, bc);
      const double alpha = rz / pq;
      // ---- phase B ----
      for (int k = tid; k < nc; k += T) sRc[k] -= alpha * sEx[KE1 * (k / CM) + 1 + (k % CM)];
      for (int k = tid; k < 3 * nloc; k += T) { sY[k] += alpha * sP[k]; sR[k] -= alpha * sQ[k]; }
      __syncthreads();
      apply_precond(a2);
      block_sum<2>(a2, red);
      __syncthreads();
      PG_TICK(tD);
      if (tid == 0) {
        __threadfence();   // z of this CTA visible before the flagged values
        double * m = e2 + (size_t)I * kSlotStride;
        m[0] = a2[0]; m[1] = a2[1];
      }
      poll_slots<2>(e2, G, sEx);
      PG_TICK(tE);
      // every CTA published E2(it) only after it finished reading E1(it): recycle own E1(it) slots
      if (tid < KE1) c.e1[((size_t)par * G + I) * kSlotStride + tid] = SENT;
      if (tid < 32) {
        double u = 0, w = 0;
        for (int i = tid; i < G; i += 32) { u += sEx[2 * i]; w += sEx[2 * i + 1]; }
        u = warp_sum(u); w = warp_sum(w);
        if (tid == 0) { s_small[0] = u; s_small[1] = w; }
      }
      __syncthreads();
      const double rz_new = s_small[0];
      rr = s_small[1];
      ++it;
      cur ^= 1;
      if (!(rr > stop) || !(pq > 0.0)) break;
      beta = rz_new / rz;
      rz = rz_new;
    }
  }
  for (int k = tid; k < 3 * nloc; k += T) d.y[3 * lo + k] = sY[k];
  if (I == 0 && tid == 0) {
    unsigned long long t_end;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
    d.scalars[8] = (double)it;
    d.scalars[9] = bb > 0.0 ? sqrt(rr / bb) : 0.0;
    d.scalars[10] = (double)(t_setup - t_start);   // ns: load rows + P~ + first barrier + Ac rows
    d.scalars[11] = (double)(t_gj - t_setup);      // ns: block Gauss-Jordan
    d.scalars[12] = (double)(t_end - t_gj);        // ns: CG iterations
    d.scalars[13] = (double)tA; d.scalars[14] = (double)tB; d.scalars[15] = (double)tC;
    d.scalars[6] = (double)tD; d.scalars[7] = (double)tE;
  }
#undef PG_TICK
}

// candidate point: delta = -(y * scale) ; xc = x (+) delta on free nodes; partial ||x - xc||^2
__global__ void __launch_bounds__(kPgThreads) k_pg_apply_step(PgDev d)
{
  __shared__ double red[32];
  double s[1] = {0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < d.N; i += gridDim.x * blockDim.x) {
    const double * x = d.x + 3 * i;
    double c0 = x[0], c1 = x[1], c2 = x[2];
    if (d.is_free[i]) {
      const double d0 = -d.y[3 * i] * d.scale[3 * i], d1 = -d.y[3 * i + 1] * d.scale[3 * i + 1],
                   d2 = -d.y[3 * i + 2] * d.scale[3 * i + 2];
      c0 = x[0] + d0; c1 = x[1] + d1; c2 = wrap_angle(x[2] + d2);   // AngleLocalParameterization, ceres_utils.h:38-54
      s[0] += (x[0] - c0) * (x[0] - c0) + (x[1] - c1) * (x[1] - c1) + (x[2] - c2) * (x[2] - c2);
    }
    d.xc[3 * i] = c0; d.xc[3 * i + 1] = c1; d.xc[3 * i + 2] = c2;
  }
  block_sum<1>(s, red);
  if (threadIdx.x == 0) d.partial[blockIdx.x] = s[0];
}

// model_cost_change = -(J~ step)^T (r + J~ step / 2) with step = -y (TrustRegionMinimizer::ComputeTrustRegionStep)
__global__ void __launch_bounds__(kPgThreads) k_pg_model_change(PgDev d)
{
  __shared__ double red[32];
  double s[1] = {0};
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < d.E; e += gridDim.x * blockDim.x) {
    const int a = d.eidx[2 * e], b = d.eidx[2 * e + 1];
    const double * L = d.lin + (size_t)kLin * e;
    const double sa0 = -d.y[3 * a], sa1 = -d.y[3 * a + 1], sa2 = -d.y[3 * a + 2];
    const double sb0 = -d.y[3 * b], sb1 = -d.y[3 * b + 1], sb2 = -d.y[3 * b + 2];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const double m = L[3 + 3 * i] * sa0 + L[4 + 3 * i] * sa1 + L[5 + 3 * i] * sa2 + L[12 + 3 * i] * sb0 +
                       L[13 + 3 * i] * sb1 + L[14 + 3 * i] * sb2;
      s[0] -= m * (L[i] + 0.5 * m);
    }
  }
  block_sum<1>(s, red);
  if (threadIdx.x == 0) d.partial[blockIdx.x] = s[0];
}

__global__ void k_pg_fill(double * p, double v, int n)
{
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = v;
}

}  // namespace b200

using namespace b200;

// ------------------------------------------------------------------------------------------
// host side: graph store + LM driver
// ------------------------------------------------------------------------------------------
struct PgEdge {
  int32_t ida, idb;
  double z[3];
  double U[6];
};

struct b200pg {
  b200pg_opts o{};
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  // graph store (mirrors CeresSolver's nodes_ / blocks_)
  std::vector<int32_t> node_ids;              // insertion order
  std::vector<double> node_pose;              // [n][3]
  std::unordered_map<int32_t, int32_t> index; // id -> position in node_ids
  std::vector<PgEdge> edges;
  int32_t first_node_id = 0;
  bool have_first = false;
  // flattened edge arrays (node positions, measurement, sqrt information), kept in step with `edges`: AddConstraint appends
  // (the mapper's normal traffic, Mapper.cpp:1634), removals / Reset mark them for a rebuild.  dev_edges of them are already
  // on the device, so a Compute after k new constraints uploads k edges, not the graph.
  std::vector<int32_t> f_eidx;
  std::vector<double> f_z, f_U;
  bool flat_dirty = false;
  size_t dev_edges = 0;
  std::vector<int32_t> agg_start_h, agg_of_h;
  // corrections of the last solve
  std::vector<int32_t> corr_ids;
  std::vector<double> corr_pose;
  // device
  DevBuf<int32_t> d_eidx, d_adj_start, d_adj, d_agg_start, d_agg_of;
  DevBuf<uint8_t> d_free;
  DevBuf<double> d_gz, d_gp, d_gPt, d_gRow, d_grc, d_e1, d_e2;
  bool debug = false;
  int precond = 1;   // 1 = two-level (rigid-mode aggregation) + block Jacobi, 0 = block Jacobi only
  int coarse_modes = 6;   // two-level: coarse modes per aggregate (6 = rigid + linear deformation, 3 = rigid only)
  DevBuf<unsigned int> d_bar;
  bool force_global_pcg = false;
  DevBuf<double> d_z, d_U, d_x, d_xc, d_scale, d_lin, d_Hd, d_g, d_diag, d_y, d_pr, d_pz, d_pp0, d_pp1, d_pq, d_Minv,
    d_partial, d_scalars;
  PinBuf<double> h_scalars;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int64_t launches = 0;
};

namespace b200 {

void b200pg_defaults(b200pg_opts * o)
{
  o->max_num_iterations = 50;
  o->function_tolerance = 1e-3;
  o->gradient_tolerance = 1e-6;
  o->parameter_tolerance = 1e-3;
  o->min_relative_decrease = 1e-3;
  o->initial_trust_region_radius = 1e4;
  o->max_trust_region_radius = 1e8;
  o->min_trust_region_radius = 1e-16;
  o->min_lm_diagonal = 1e-6;
  o->max_lm_diagonal = 1e32;
  o->jacobi_scaling = 1;
  o->use_nonmonotonic_steps = 1;
  o->max_consecutive_nonmonotonic_steps = 3;
  o->max_num_consecutive_invalid_steps = 3;
  o->pcg_tolerance = 1e-9;    // loosest tolerance that keeps cfg4 (0.05 m / 0.02 rad) within 5e-6 m of the exact-solve LM (tools/pcg_tolerance_study.py)
  o->pcg_max_iterations = 20000;
  o->loss_function = 0;
  o->loss_scale = 0.7;
}

// karto::Matrix3::Inverse by cofactors (Karto.h:2533-2577) then Eigen's llt().matrixU()
// of the matrix rebuilt from its upper triangle (ceres_solver.cpp:364-376)
static bool sqrt_information(const double cov[9], double U[6])
{
  const double * m = cov;
  double inv[9];
  inv[0] = m[4] * m[8] - m[5] * m[7];
  inv[1] = m[2] * m[7] - m[1] * m[8];
  inv[2] = m[1] * m[5] - m[2] * m[4];
  inv[3] = m[5] * m[6] - m[3] * m[8];
  inv[4] = m[0] * m[8] - m[2] * m[6];
  inv[5] = m[2] * m[3] - m[0] * m[5];
  inv[6] = m[3] * m[7] - m[4] * m[6];
  inv[7] = m[1] * m[6] - m[0] * m[7];
  inv[8] = m[0] * m[4] - m[1] * m[3];
  double det = m[0] * inv[0] + m[1] * inv[3] + m[2] * inv[6];
  if (!(fabs(det) > 1e-14)) return false;   // Matrix3::Inverse asserts
  double id = 1.0 / det;
  for (int i = 0; i < 9; ++i) inv[i] *= id;
  const double a00 = inv[0], a01 = inv[1], a02 = inv[2], a11 = inv[4], a12 = inv[5], a22 = inv[8];
  if (!(a00 > 0)) return false;
  const double l00 = sqrt(a00), l10 = a01 / l00, l20 = a02 / l00;
  const double t11 = a11 - l10 * l10;
  if (!(t11 > 0)) return false;
  const double l11 = sqrt(t11), l21 = (a12 - l20 * l10) / l11;
  const double t22 = a22 - l20 * l20 - l21 * l21;
  if (!(t22 > 0)) return false;
  const double l22 = sqrt(t22);
  U[0] = l00; U[1] = l10; U[2] = l20; U[3] = l11; U[4] = l21; U[5] = l22;
  return true;
}

template <class T>
static void up(DevBuf<T> & dst, const std::vector<T> & src, cudaStream_t s)
{
  dst.reserve(std::max<size_t>(src.size(), 1));
  if (!src.empty()) B200_CUDA(cudaMemcpyAsync(dst.p, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice, s));
}

// grow a device buffer to n elements keeping its first `keep` ones
template <class T>
static void grow_keep(DevBuf<T> & b, size_t n, size_t keep, cudaStream_t s)
{
  if (n <= b.cap) return;
  if (keep == 0 || !b.p) { b.reserve(n); return; }
  DevBuf<T> nb;
  nb.reserve(n + n / 2);
  B200_CUDA(cudaMemcpyAsync(nb.p, b.p, keep * sizeof(T), cudaMemcpyDeviceToDevice, s));
  B200_CUDA(cudaStreamSynchronize(s));
  std::swap(b.p, nb.p); std::swap(b.cap, nb.cap);
}

struct Lm {
  b200pg * h;
  PgDev d;
  int blocksN, blocksE, pcg_blocks;
  bool use_smem = false;
  int smem_blocks = 0;
  size_t smem_bytes = 0;
  PcgSmemCfg cfg{};
  bool use_2lvl = false;
  int blocks2 = 0;
  size_t smem2_bytes = 0;
  int cm2 = 3;
  Pcg2Cfg cfg2{};
  cudaStream_t st;

  double scalar(int slot)
  {
    B200_CUDA(cudaMemcpyAsync(h->h_scalars.p, h->d_scalars.p, 16 * sizeof(double), cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
    return h->h_scalars.p[slot];
  }
  void fetch()
  {
    B200_CUDA(cudaMemcpyAsync(h->h_scalars.p, h->d_scalars.p, 16 * sizeof(double), cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
  }
  void launched() { B200_CUDA(cudaGetLastError()); h->launches++; }

  // cost at x (mode 0, also linearises) or at xc (mode 1) -> scalars[slot]
  void linearize(int mode, int slot)
  {
    k_pg_linearize<<<blocksE, kPgThreads, 0, st>>>(d, mode); launched();
    k_pg_reduce<<<1, 32, 0, st>>>(d, blocksE, slot, -1); launched();
  }
  // Hd, g, gradient max norm (scalars[2]) and ||x||^2 (scalars[3])
  void assemble()
  {
    k_pg_assemble<<<blocksN, kPgThreads, 0, st>>>(d, 0); launched();
    k_pg_reduce<<<1, 32, 0, st>>>(d, blocksN, 3, 2); launched();
  }
};

static int solve(b200pg * h, b200pg_summary * sum)
{
  const auto t_enter = std::chrono::steady_clock::now();
  NvtxRange nvtx_solve("b200pg solve");
  const b200pg_opts & o = h->o;
  b200pg_summary S{};
  S.usable = 1;
  const int N = (int)h->node_ids.size();
  if (N == 0) {   // "Ceres was called when there are no nodes" (ceres_solver.cpp:219-225)
    set_last_error("b200pg_solve: no nodes");
    if (sum) *sum = S;
    return B200_ERR_INVALID_ARG;
  }
  require_device();
  if (!h->stream) { B200_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)); h->own_stream = true; }
  cudaStream_t st = h->stream;
  if (!h->ev0) { B200_CUDA(cudaEventCreate(&h->ev0)); B200_CUDA(cudaEventCreate(&h->ev1)); }
  const int64_t launches0 = h->launches;

  // ---- flattened graph: kept incrementally (see b200pg::f_eidx) ----
  const int E = (int)h->edges.size();
  if (h->flat_dirty || h->f_eidx.size() != 2 * (size_t)E) {
    h->f_eidx.resize(2 * (size_t)E); h->f_z.resize(3 * (size_t)E); h->f_U.resize(6 * (size_t)E);
    for (int e = 0; e < E; ++e) {
      const PgEdge & ed = h->edges[e];
      h->f_eidx[2 * e] = h->index.at(ed.ida); h->f_eidx[2 * e + 1] = h->index.at(ed.idb);
      for (int k = 0; k < 3; ++k) h->f_z[3 * e + k] = ed.z[k];
      for (int k = 0; k < 6; ++k) h->f_U[6 * e + k] = ed.U[k];
    }
    h->flat_dirty = false;
    h->dev_edges = 0;
  }
  const std::vector<int32_t> & eidx = h->f_eidx;
  std::vector<uint8_t> is_free(N, 0);
  std::vector<int32_t> deg(N + 1, 0);
  for (int e = 0; e < E; ++e) {
    const int a = eidx[2 * e], b = eidx[2 * e + 1];
    is_free[a] = 1; is_free[b] = 1;   // only nodes that appear in a residual block are Ceres parameter blocks
    deg[a + 1]++; deg[b + 1]++;
  }
  // first node constant, if it is part of the problem (ceres_solver.cpp:228-241)
  if (h->have_first) {
    auto it = h->index.find(h->first_node_id);
    if (it != h->index.end()) is_free[it->second] = 0;
  }
  int nfree = 0;
  for (int i = 0; i < N; ++i) nfree += is_free[i];
  auto store_corrections = [&]() {
    h->corr_ids = h->node_ids;
    h->corr_pose = h->node_pose;
  };
  if (E == 0 || nfree == 0) {   // nothing to optimise: Ceres returns CONVERGENCE immediately
    store_corrections();
    S.termination = 0;
    if (sum) *sum = S;
    return B200_OK;
  }
  std::vector<int32_t> adj_start(N + 1, 0), adj(2 * (size_t)E);
  for (int i = 0; i < N; ++i) adj_start[i + 1] = adj_start[i] + deg[i + 1];
  {
    std::vector<int32_t> fill(adj_start.begin(), adj_start.end() - 1);
    for (int e = 0; e < E; ++e) {
      adj[fill[eidx[2 * e]]++] = (e << 1) | 0;
      adj[fill[eidx[2 * e + 1]]++] = (e << 1) | 1;
    }
  }
  {
    // edges: only the ones appended since the last solve travel (the buffers grow with their contents kept)
    if (h->dev_edges > (size_t)E) h->dev_edges = 0;
    const size_t e0 = h->dev_edges, ne = (size_t)E - e0;
    grow_keep(h->d_eidx, 2 * (size_t)E, 2 * e0, st); grow_keep(h->d_z, 3 * (size_t)E, 3 * e0, st); grow_keep(h->d_U, 6 * (size_t)E, 6 * e0, st);
    if (ne) {
      B200_CUDA(cudaMemcpyAsync(h->d_eidx.p + 2 * e0, h->f_eidx.data() + 2 * e0, 2 * ne * sizeof(int32_t), cudaMemcpyHostToDevice, st));
      B200_CUDA(cudaMemcpyAsync(h->d_z.p + 3 * e0, h->f_z.data() + 3 * e0, 3 * ne * sizeof(double), cudaMemcpyHostToDevice, st));
      B200_CUDA(cudaMemcpyAsync(h->d_U.p + 6 * e0, h->f_U.data() + 6 * e0, 6 * ne * sizeof(double), cudaMemcpyHostToDevice, st));
    }
    S.uploaded_edges = (int32_t)ne;
    h->dev_edges = (size_t)E;
  }
  up(h->d_free, is_free, st);
  up(h->d_adj_start, adj_start, st); up(h->d_adj, adj, st); up(h->d_x, h->node_pose, st);
  const size_t n3 = 3 * (size_t)N;
  h->d_xc.reserve(n3); h->d_scale.reserve(n3); h->d_lin.reserve((size_t)kLin * E); h->d_Hd.reserve(6 * (size_t)N);
  h->d_g.reserve(n3); h->d_diag.reserve(n3); h->d_y.reserve(n3); h->d_pr.reserve(n3); h->d_pz.reserve(n3);
  h->d_pp0.reserve(n3); h->d_pp1.reserve(n3); h->d_pq.reserve(n3); h->d_Minv.reserve(6 * (size_t)N);
  h->d_partial.reserve((size_t)kMaxPartials * 8); h->d_scalars.reserve(16); h->h_scalars.reserve(16);

  Lm L;
  L.h = h; L.st = st;
  PgDev & d = L.d;
  d.N = N; d.E = E; d.eidx = h->d_eidx.p; d.z = h->d_z.p; d.U = h->d_U.p; d.is_free = h->d_free.p;
  d.adj_start = h->d_adj_start.p; d.adj = h->d_adj.p; d.x = h->d_x.p; d.xc = h->d_xc.p; d.scale = h->d_scale.p;
  d.lin = h->d_lin.p; d.Hd = h->d_Hd.p; d.g = h->d_g.p; d.diag = h->d_diag.p; d.y = h->d_y.p; d.pr = h->d_pr.p;
  d.pz = h->d_pz.p; d.pp0 = h->d_pp0.p; d.pp1 = h->d_pp1.p; d.pq = h->d_pq.p; d.Minv = h->d_Minv.p;
  d.partial = h->d_partial.p; d.scalars = h->d_scalars.p;
  d.loss = o.loss_function; d.loss_a = o.loss_scale;
  int dev = 0, sms = 148, per_sm = 1;
  B200_CUDA(cudaGetDevice(&dev));
  B200_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pg_pcg, kPgThreads, 0));
  L.blocksN = std::min(kMaxPartials, std::max(1, (N + kPgThreads - 1) / kPgThreads));
  L.blocksE = std::min(kMaxPartials, std::max(1, (E + kPgThreads - 1) / kPgThreads));
  L.pcg_blocks = std::max(1, std::min({kMaxPartials, sms * std::max(per_sm, 1), (N + kPgThreads - 1) / kPgThreads}));

  {
    // shared-memory-resident PCG when every CTA's rows fit (one CTA per SM, 512 threads)
    const int G = std::min(sms, std::max(1, (N + 31) / 32));
    const int npc = (N + G - 1) / G;
    const int Gu = (N + npc - 1) / npc;
    int max_slots = 0;
    for (int c = 0; c < Gu; ++c) {
      const int lo = c * npc, hi = std::min(N, lo + npc);
      max_slots = std::max(max_slots, adj_start[hi] - adj_start[lo]);
    }
    max_slots = std::max(max_slots, 1);
    const size_t bytes = ((size_t)max_slots * 12 + (size_t)npc * 27) * sizeof(double) + ((size_t)max_slots + npc + 1) * sizeof(int) + 16;
    if (bytes <= 200 * 1024 && !h->force_global_pcg) {
      L.use_smem = true; L.smem_blocks = Gu; L.smem_bytes = bytes;
      h->d_gz.reserve(n3); h->d_gp.reserve(2 * n3); h->d_bar.reserve(4096);
      L.cfg.npc = npc; L.cfg.max_slots = max_slots; L.cfg.gz = h->d_gz.p; L.cfg.gp = h->d_gp.p; L.cfg.bar = h->d_bar.p;
      B200_CUDA(cudaFuncSetAttribute(k_pg_pcg_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    }
  }
  if (h->precond == 1 && !h->force_global_pcg) {
    // two-level preconditioner: one aggregate per 256-thread CTA. One CTA per SM is preferred: with two per SM
    // (295 aggregates at cfg4) CG needs 28 % fewer iterations, but the all-to-all exchanges and the 295-step
    // Gauss-Jordan get slower (20 us/iteration instead of 12.6; measured 120 ms vs 95 ms). Two per SM is the
    // fallback when one aggregate per SM does not fit shared memory.
    for (int per_sm = 1; per_sm <= 2 && !L.use_2lvl; ++per_sm) {
      const int G2 = std::min(per_sm * sms, std::max(1, (N + 15) / 16));
      // contiguous node ranges of equal node count: compact aggregates give the best coarse space (balancing by
      // block count was tried: it merges sparse chain stretches into large aggregates and costs 60 % more iterations)
      std::vector<int32_t> & agg_start = h->agg_start_h;
      std::vector<int32_t> & agg_of = h->agg_of_h;
      agg_start.clear(); agg_of.assign(N, 0);
      {
        const int per = (N + G2 - 1) / G2;
        for (int i = 0; i < N; i += per) agg_start.push_back(i);
        agg_start.push_back(N);
        for (int a = 0; a + 1 < (int)agg_start.size(); ++a)
          for (int i = agg_start[a]; i < agg_start[a + 1]; ++i) agg_of[i] = a;
      }
      const int Gu = (int)agg_start.size() - 1;
      int max_slots = 1, npc = 1;
      for (int c = 0; c < Gu; ++c) {
        max_slots = std::max(max_slots, adj_start[agg_start[c + 1]] - adj_start[agg_start[c]]);
        npc = std::max(npc, agg_start[c + 1] - agg_start[c]);
      }
      // coarse modes per aggregate: 6 (rigid + piecewise-linear deformation) when it fits shared memory, else 3
      int cm = 0, nc = 0, ex_doubles = 0;
      size_t bytes2 = 0;
      for (int try_cm : {6, 3}) {
        if (try_cm > h->coarse_modes) continue;
        nc = try_cm * Gu;
        ex_doubles = std::max((1 + try_cm) * Gu, try_cm * nc - 3 * max_slots);
        bytes2 = ((size_t)max_slots * 12 + (size_t)npc * 37 + (size_t)(try_cm + 1) * nc + (size_t)ex_doubles) * sizeof(double) +
                 ((size_t)2 * max_slots + npc + 1) * sizeof(int) + 16;
        if (bytes2 > 220 * 1024) continue;
        int occ = 0;
        const void * fn = try_cm == 6 ? (const void *)k_pg_pcg_2lvl<6> : (const void *)k_pg_pcg_2lvl<3>;
        B200_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes2));
        B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, 256, bytes2));
        if (h->debug) fprintf(stderr, "[b200pg] two-level plan: %d aggregates x %d modes, <= %d nodes and <= %d blocks each, %zu B smem, occupancy %d/SM\n", Gu, try_cm, npc, max_slots, bytes2, occ);
        if (occ * sms < Gu) continue;
        cm = try_cm;
        break;
      }
      if (!cm) continue;
      L.use_2lvl = true; L.smem2_bytes = bytes2; L.blocks2 = Gu; L.cm2 = cm;
      h->d_gz.reserve(n3); h->d_gp.reserve(2 * n3);
      h->d_bar.reserve(std::max<size_t>(4096, (size_t)Gu + 4));
      h->d_gPt.reserve(10 * (size_t)N); h->d_gRow.reserve((size_t)Gu * cm * 2 * nc); h->d_grc.reserve(nc);
      h->d_e1.reserve((size_t)2 * Gu * kSlotStride); h->d_e2.reserve((size_t)2 * Gu * kSlotStride);
      up(h->d_agg_start, agg_start, st); up(h->d_agg_of, agg_of, st);   // the vectors live in the handle
      Pcg2Cfg & c2 = L.cfg2;
      c2.npc = npc; c2.max_slots = max_slots; c2.ex_doubles = ex_doubles; c2.agg_start = h->d_agg_start.p; c2.agg_of = h->d_agg_of.p;
      c2.gz = h->d_gz.p; c2.gp = h->d_gp.p; c2.gPt = h->d_gPt.p; c2.gRow = h->d_gRow.p;
      c2.grc = h->d_grc.p; c2.e1 = h->d_e1.p; c2.e2 = h->d_e2.p; c2.bar = h->d_bar.p; c2.gjflag = h->d_bar.p + 1;
    }
  }
  S.setup_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_enter).count();
  B200_CUDA(cudaEventRecord(h->ev0, st));
  // ---- iteration 0: evaluate, Jacobi scaling from the unscaled Jacobian ----
  k_pg_fill<<<64, 256, 0, st>>>(d.scale, 1.0, 3 * N); L.launched();
  if (o.jacobi_scaling) {
    L.linearize(0, 0);
    k_pg_assemble<<<L.blocksN, kPgThreads, 0, st>>>(d, 1); L.launched();
  }
  L.linearize(0, 0);   // scalars[0] = cost(x)
  L.assemble();        // scalars[2] = gradient max norm, scalars[3] = ||x||^2
  L.fetch();
  double cost = h->h_scalars.p[0], gmax = h->h_scalars.p[2], x_norm = sqrt(h->h_scalars.p[3]);
  S.initial_cost = cost;
  double minimum_cost = cost;
  std::vector<double> best_x;   // empty = the start point
  bool have_best_on_device_x = true;

  double radius = o.initial_trust_region_radius, decrease_factor = 2.0;
  bool reuse_diagonal = false;
  const int max_nonmono = o.use_nonmonotonic_steps ? o.max_consecutive_nonmonotonic_steps : 0;
  double ev_min = cost, ev_cur = cost, ev_ref = cost, ev_cand = cost, acc_ref = 0.0, acc_cand = 0.0;
  int n_nonmono = 0, invalid_steps = 0, it = 0;
  bool step_successful = false;
  S.termination = 3;
  // TrustRegionMinimizer::IterationZero: an already-converged start returns CONVERGENCE before any step is computed
  const bool converged_at_start = gmax <= o.gradient_tolerance;
  if (converged_at_start) S.termination = 1;
  while (!converged_at_start) {
    // FinalizeIterationAndCheckIfMinimizerCanContinue
    if (it >= o.max_num_iterations) { S.termination = 3; break; }
    if (step_successful && gmax <= o.gradient_tolerance) { S.termination = 1; break; }
    if (radius <= o.min_trust_region_radius) { S.termination = 4; break; }
    ++it;
    NvtxRange nvtx_it("b200pg LM iteration");
    step_successful = false;
    // LevenbergMarquardtStrategy::ComputeStep
    if (!reuse_diagonal) { k_pg_diag<<<L.blocksN, kPgThreads, 0, st>>>(d, o.min_lm_diagonal, o.max_lm_diagonal); L.launched(); }
    {
      double inv_radius = 1.0 / radius, tol = o.pcg_tolerance;
      int max_iter = o.pcg_max_iterations;
      if (L.use_2lvl) {
        B200_CUDA(cudaMemsetAsync(h->d_bar.p, 0, (size_t)(L.blocks2 + 1) * sizeof(unsigned int), st));
        void * args[] = {&d, &L.cfg2, &inv_radius, &tol, &max_iter};
        B200_CUDA(cudaLaunchCooperativeKernel(L.cm2 == 6 ? (void *)k_pg_pcg_2lvl<6> : (void *)k_pg_pcg_2lvl<3>, dim3(L.blocks2), dim3(256), args, L.smem2_bytes, st));
      } else if (L.use_smem) {
        B200_CUDA(cudaMemsetAsync(h->d_bar.p, 0, sizeof(unsigned int), st));
        void * args[] = {&d, &L.cfg, &inv_radius, &tol, &max_iter};
        B200_CUDA(cudaLaunchCooperativeKernel((void *)k_pg_pcg_smem, dim3(L.smem_blocks), dim3(512), args, L.smem_bytes, st));
      } else {
        void * args[] = {&d, &inv_radius, &tol, &max_iter};
        B200_CUDA(cudaLaunchCooperativeKernel((void *)k_pg_pcg, dim3(L.pcg_blocks), dim3(kPgThreads), args, 0, st));
      }
      h->launches++;
    }
    reuse_diagonal = true;
    k_pg_model_change<<<L.blocksE, kPgThreads, 0, st>>>(d); L.launched();
    k_pg_reduce<<<1, 32, 0, st>>>(d, L.blocksE, 4, -1); L.launched();
    k_pg_apply_step<<<L.blocksN, kPgThreads, 0, st>>>(d); L.launched();
    k_pg_reduce<<<1, 32, 0, st>>>(d, L.blocksN, 5, -1); L.launched();
    L.linearize(1, 1);   // scalars[1] = cost(xc)
    L.fetch();
    const double * sc = h->h_scalars.p;
    S.pcg_iterations += (int)sc[8];
    if (h->debug && L.use_2lvl)
      fprintf(stderr, "[b200pg] lm %d: pcg %d it, setup %.1f us, gauss-jordan %.1f us, cg %.1f us (%.2f us/it: gather %.2f spmv+reduce %.2f exch1 %.2f precond %.2f exch2 %.2f)\n", it, (int)sc[8],
              sc[10] * 1e-3, sc[11] * 1e-3, sc[12] * 1e-3, sc[12] * 1e-3 / std::max(1.0, sc[8]), sc[13] * 1e-3 / std::max(1.0, sc[8]),
              sc[14] * 1e-3 / std::max(1.0, sc[8]), sc[15] * 1e-3 / std::max(1.0, sc[8]), sc[6] * 1e-3 / std::max(1.0, sc[8]), sc[7] * 1e-3 / std::max(1.0, sc[8]));
    const double model_cost_change = sc[4], step_norm = sqrt(sc[5]), cand_cost = sc[1];
    const bool finite = std::isfinite(model_cost_change) && std::isfinite(cand_cost) && std::isfinite(sc[9]);
    const bool valid = finite && model_cost_change > 0.0;
    if (!valid) {   // HandleInvalidStep
      if (++invalid_steps >= o.max_num_consecutive_invalid_steps) { S.termination = 5; S.usable = 0; break; }
      radius /= decrease_factor; decrease_factor *= 2.0;
      continue;
    }
    invalid_steps = 0;
    if (step_norm <= o.parameter_tolerance * (x_norm + o.parameter_tolerance)) { S.termination = 2; break; }
    const double cost_change = cost - cand_cost;
    if (fabs(cost_change) <= o.function_tolerance * cost) { S.termination = 0; break; }
    const double rel = (ev_cur - cand_cost) / model_cost_change;
    const double hist = (ev_ref - cand_cost) / (acc_ref + model_cost_change);
    const double quality = std::max(rel, hist);
    if (quality > o.min_relative_decrease) {   // HandleSuccessfulStep
      std::swap(d.x, d.xc);
      cost = cand_cost;
      L.linearize(0, 0);
      L.assemble();
      L.fetch();
      gmax = h->h_scalars.p[2]; x_norm = sqrt(h->h_scalars.p[3]);
      step_successful = true;
      S.successful_steps++;
      radius = std::min(o.max_trust_region_radius, radius / std::max(1.0 / 3.0, 1.0 - pow(2.0 * quality - 1.0, 3)));
      decrease_factor = 2.0; reuse_diagonal = false;
      ev_cur = cost; acc_cand += model_cost_change; acc_ref += model_cost_change;
      if (ev_cur < ev_min) { ev_min = ev_cur; n_nonmono = 0; ev_cand = ev_cur; acc_cand = 0.0; }
      else { ++n_nonmono; if (ev_cur > ev_cand) { ev_cand = ev_cur; acc_cand = 0.0; } }
      if (n_nonmono == max_nonmono) { ev_ref = ev_cand; acc_ref = acc_cand; }
      if (cost < minimum_cost) {   // the solution Ceres returns is the minimum-cost iterate
        minimum_cost = cost;
        have_best_on_device_x = true;
      } else if (have_best_on_device_x) {
        // x moved away from the best iterate (non-monotonic step): keep a host copy of the best one,
        // which is the previous x, now in d.xc
        best_x.resize(n3);
        B200_CUDA(cudaMemcpyAsync(best_x.data(), d.xc, n3 * sizeof(double), cudaMemcpyDeviceToHost, st));
        B200_CUDA(cudaStreamSynchronize(st));
        have_best_on_device_x = false;
      }
    } else {   // HandleUnsuccessfulStep
      radius /= decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;
    }
  }
  B200_CUDA(cudaEventRecord(h->ev1, st));
  S.iterations = it;
  S.final_cost = minimum_cost;
  if (S.usable) {
    if (have_best_on_device_x) {
      B200_CUDA(cudaMemcpyAsync(h->node_pose.data(), d.x, n3 * sizeof(double), cudaMemcpyDeviceToHost, st));
    } else {
      std::memcpy(h->node_pose.data(), best_x.data(), n3 * sizeof(double));
    }
  }
  B200_CUDA(cudaStreamSynchronize(st));
  B200_CUDA(cudaEventElapsedTime(&S.solve_ms, h->ev0, h->ev1));
  S.kernel_launches = h->launches - launches0;
  S.wall_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_enter).count();
  if (sum) *sum = S;
  if (!S.usable) {
    set_last_error("pose-graph solve produced no usable solution (too many invalid steps)");
    return B200_ERR_NUMERIC;
  }
  store_corrections();
  return B200_OK;
}

}  // namespace b200

#define B200_GUARD_BEGIN try {
#define B200_GUARD_END                                                     \
  }                                                                        \
  catch (const b200::CudaFail & f) { return f.code; }                      \
  catch (const std::bad_alloc &) { b200::set_last_error("out of host memory"); return B200_ERR_CUDA; } \
  catch (const std::exception & e) { b200::set_last_error(e.what()); return B200_ERR_CUDA; }

extern "C" {

void b200pg_default_opts(b200pg_opts * o) { if (o) b200pg_defaults(o); }

int b200pg_create(const b200pg_opts * opts, b200pg ** out)
{
  B200_GUARD_BEGIN
  if (!out) return B200_ERR_INVALID_ARG;
  *out = nullptr;
  std::unique_ptr<b200pg> h(new b200pg());
  if (opts) h->o = *opts; else b200pg_defaults(&h->o);
  if (h->o.max_num_iterations < 0 || !(h->o.pcg_tolerance > 0) || h->o.pcg_max_iterations <= 0 ||
      !(h->o.initial_trust_region_radius > 0) || h->o.loss_function < 0 || h->o.loss_function > 2 ||
      (h->o.loss_function != 0 && !(h->o.loss_scale > 0))) {
    set_last_error("b200pg_create: invalid options");
    return B200_ERR_INVALID_ARG;
  }
  require_device();
  if (const char * e = getenv("B200PG_FORCE_GLOBAL_PCG")) h->force_global_pcg = atoi(e) != 0;
  if (const char * e = getenv("B200PG_DEBUG")) h->debug = atoi(e) != 0;
  if (const char * e = getenv("B200PG_PRECOND")) h->precond = std::string(e) == "jacobi" ? 0 : 1;
  if (const char * e = getenv("B200PG_COARSE_MODES")) h->coarse_modes = atoi(e) >= 6 ? 6 : 3;
  *out = h.release();
  return B200_OK;
  B200_GUARD_END
}

void b200pg_destroy(b200pg * h)
{
  if (!h) return;
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

static bool valid_opts(const b200pg_opts & o)
{
  return !(o.max_num_iterations < 0 || !(o.pcg_tolerance > 0) || o.pcg_max_iterations <= 0 || !(o.initial_trust_region_radius > 0) ||
           o.loss_function < 0 || o.loss_function > 2 || (o.loss_function != 0 && !(o.loss_scale > 0)));
}

int b200pg_set_opts(b200pg * h, const b200pg_opts * opts)
{
  if (!h || !opts) return B200_ERR_INVALID_ARG;
  if (!valid_opts(*opts)) { set_last_error("b200pg_set_opts: invalid options"); return B200_ERR_INVALID_ARG; }
  h->o = *opts;
  return B200_OK;
}

int b200pg_get_opts(const b200pg * h, b200pg_opts * opts)
{
  if (!h || !opts) return B200_ERR_INVALID_ARG;
  *opts = h->o;
  return B200_OK;
}

int b200pg_set_stream(b200pg * h, void * s)
{
  if (!h) return B200_ERR_INVALID_ARG;
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  h->stream = static_cast<cudaStream_t>(s);
  h->own_stream = false;
  return B200_OK;
}

int b200pg_reset(b200pg * h)
{
  if (!h) return B200_ERR_INVALID_ARG;
  h->node_ids.clear(); h->node_pose.clear(); h->index.clear(); h->edges.clear();
  h->corr_ids.clear(); h->corr_pose.clear();
  h->have_first = false;
  h->f_eidx.clear(); h->f_z.clear(); h->f_U.clear(); h->flat_dirty = false; h->dev_edges = 0;
  return B200_OK;
}

int b200pg_clear(b200pg * h)
{
  if (!h) return B200_ERR_INVALID_ARG;
  h->corr_ids.clear(); h->corr_pose.clear();
  return B200_OK;
}

int b200pg_add_node(b200pg * h, int32_t id, const double pose[3])
{
  if (!h || !pose) return B200_ERR_INVALID_ARG;
  if (h->index.count(id)) return B200_OK;   // unordered_map::insert keeps the existing entry (ceres_solver.cpp:331)
  h->index[id] = (int32_t)h->node_ids.size();
  h->node_ids.push_back(id);
  h->node_pose.insert(h->node_pose.end(), pose, pose + 3);
  if (h->node_ids.size() == 1) { h->first_node_id = id; h->have_first = true; }   // :333-335
  return B200_OK;
}

int b200pg_add_edge(b200pg * h, int32_t ida, int32_t idb, const double z[3], const double cov[9])
{
  if (!h || !z || !cov) return B200_ERR_INVALID_ARG;
  if (!h->index.count(ida) || !h->index.count(idb) || ida == idb) {
    set_last_error("b200pg_add_edge: could not find nodes (CeresSolver warns and ignores, ceres_solver.cpp:351-358)");
    return B200_ERR_NOT_FOUND;
  }
  PgEdge e;
  e.ida = ida; e.idb = idb;
  e.z[0] = z[0]; e.z[1] = z[1]; e.z[2] = z[2];
  if (!sqrt_information(cov, e.U)) {
    set_last_error("b200pg_add_edge: covariance is singular or its inverse is not positive definite");
    return B200_ERR_NUMERIC;
  }
  h->edges.push_back(e);
  if (!h->flat_dirty) {
    h->f_eidx.push_back(h->index[ida]); h->f_eidx.push_back(h->index[idb]);
    h->f_z.insert(h->f_z.end(), e.z, e.z + 3);
    h->f_U.insert(h->f_U.end(), e.U, e.U + 6);
  }
  return B200_OK;
}

int b200pg_remove_node(b200pg * h, int32_t id)
{
  if (!h) return B200_ERR_INVALID_ARG;
  auto it = h->index.find(id);
  if (it == h->index.end()) { set_last_error("RemoveNode: failed to find node"); return B200_ERR_NOT_FOUND; }
  // RemoveParameterBlock also removes every residual block that uses it (ceres_solver.cpp:405-407)
  h->edges.erase(std::remove_if(h->edges.begin(), h->edges.end(), [&](const PgEdge & e) { return e.ida == id || e.idb == id; }),
                 h->edges.end());
  const int pos = it->second, last = (int)h->node_ids.size() - 1;
  h->flat_dirty = true;   // node positions move and edges go: the flattened arrays are rebuilt by the next solve
  h->index.erase(it);
  if (pos != last) {
    h->node_ids[pos] = h->node_ids[last];
    for (int k = 0; k < 3; ++k) h->node_pose[3 * pos + k] = h->node_pose[3 * last + k];
    h->index[h->node_ids[pos]] = pos;
  }
  h->node_ids.pop_back();
  h->node_pose.resize(3 * h->node_ids.size());
  return B200_OK;
}

int b200pg_remove_edge(b200pg * h, int32_t ida, int32_t idb)
{
  if (!h) return B200_ERR_INVALID_ARG;
  // the block stored under (source,target) first, else (target,source) (ceres_solver.cpp:432-447)
  for (int pass = 0; pass < 2; ++pass) {
    for (size_t k = 0; k < h->edges.size(); ++k) {
      const PgEdge & e = h->edges[k];
      if ((pass == 0 && e.ida == ida && e.idb == idb) || (pass == 1 && e.ida == idb && e.idb == ida)) {
        h->edges.erase(h->edges.begin() + k);
        h->flat_dirty = true;
        return B200_OK;
      }
    }
  }
  set_last_error("RemoveConstraint: failed to find residual block");
  return B200_ERR_NOT_FOUND;
}

int b200pg_modify_node(b200pg * h, int32_t id, const double pose[3])
{
  if (!h || !pose) return B200_ERR_INVALID_ARG;
  auto it = h->index.find(id);
  if (it == h->index.end()) return B200_ERR_NOT_FOUND;
  double * p = &h->node_pose[3 * it->second];
  const double yaw_init = p[2];
  p[0] = pose[0]; p[1] = pose[1]; p[2] = pose[2];
  p[2] += yaw_init;   // ceres_solver.cpp:457-459
  return B200_OK;
}

int b200pg_get_node(const b200pg * h, int32_t id, double pose[3])
{
  if (!h || !pose) return B200_ERR_INVALID_ARG;
  auto it = h->index.find(id);
  if (it == h->index.end()) return B200_ERR_NOT_FOUND;
  for (int k = 0; k < 3; ++k) pose[k] = h->node_pose[3 * it->second + k];
  return B200_OK;
}

int32_t b200pg_num_nodes(const b200pg * h) { return h ? (int32_t)h->node_ids.size() : 0; }
int32_t b200pg_num_edges(const b200pg * h) { return h ? (int32_t)h->edges.size() : 0; }

int b200pg_solve(b200pg * h, b200pg_summary * summary)
{
  B200_GUARD_BEGIN
  if (!h) return B200_ERR_INVALID_ARG;
  return solve(h, summary);
  B200_GUARD_END
}

int32_t b200pg_get_corrections(const b200pg * h, int32_t * ids, double * poses, int32_t cap)
{
  if (!h || !ids || !poses) return 0;
  int32_t n = (int32_t)std::min<size_t>(h->corr_ids.size(), (size_t)std::max(cap, 0));
  for (int32_t i = 0; i < n; ++i) {
    ids[i] = h->corr_ids[i];
    for (int k = 0; k < 3; ++k) poses[3 * i + k] = h->corr_pose[3 * i + k];
  }
  return n;
}

}  // extern "C"
