"""ORACLE (test infrastructure) -- CPU restatement of the SE(2) pose-graph solve behind
karto::ScanSolver as solver_plugins::CeresSolver performs it.

PARITY UNPINNED: the arithmetic of the reference lives in Ceres Solver (third-party, version
unpinned: package.xml:32 `libceres-dev`, pre-2.2 API), which is absent from /root/reference and
cannot be built here (no Eigen / SuiteSparse / glog).  The reference holds no test, golden
vector or fixture for this boundary either.  This file therefore restates
  * the problem the reference builds: solvers/ceres_utils.h:27-32 (angle wrap), :84-100
    (residual), solvers/ceres_solver.cpp:364-376 (covariance -> information -> upper Cholesky
    factor), :228-241 (first node constant), :158-186 (options);
  * Ceres' published trust-region Levenberg-Marquardt loop (TrustRegionMinimizer,
    LevenbergMarquardtStrategy, TrustRegionStepEvaluator) with an exact sparse solve
    (SciPy SuperLU) in place of SPARSE_NORMAL_CHOLESKY;
and is cross-checked against scipy.optimize.least_squares in tests/test_posegraph_oracle.py.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


def normalize_angle(a):
    """solvers/ceres_utils.h:27-32: [-pi, pi)."""
    two_pi = 2.0 * math.pi
    return a - two_pi * np.floor((a + math.pi) / two_pi)


def matrix3_inverse(m: np.ndarray) -> np.ndarray:
    """karto::Matrix3::Inverse by cofactors (Karto.h:2533-2577)."""
    inv = np.empty((3, 3))
    inv[0, 0] = m[1, 1] * m[2, 2] - m[1, 2] * m[2, 1]
    inv[0, 1] = m[0, 2] * m[2, 1] - m[0, 1] * m[2, 2]
    inv[0, 2] = m[0, 1] * m[1, 2] - m[0, 2] * m[1, 1]
    inv[1, 0] = m[1, 2] * m[2, 0] - m[1, 0] * m[2, 2]
    inv[1, 1] = m[0, 0] * m[2, 2] - m[0, 2] * m[2, 0]
    inv[1, 2] = m[0, 2] * m[1, 0] - m[0, 0] * m[1, 2]
    inv[2, 0] = m[1, 0] * m[2, 1] - m[1, 1] * m[2, 0]
    inv[2, 1] = m[0, 1] * m[2, 0] - m[0, 0] * m[2, 1]
    inv[2, 2] = m[0, 0] * m[1, 1] - m[0, 1] * m[1, 0]
    det = m[0, 0] * inv[0, 0] + m[0, 1] * inv[1, 0] + m[0, 2] * inv[2, 0]
    if abs(det) <= 1e-14:
        raise np.linalg.LinAlgError("singular covariance (Matrix3::Inverse asserts)")
    return inv * (1.0 / det)


def sqrt_information(cov: np.ndarray) -> np.ndarray:
    """solvers/ceres_solver.cpp:364-376: information from the upper triangle of cov^-1, U = llt().matrixU()."""
    p = matrix3_inverse(np.asarray(cov, dtype=np.float64).reshape(3, 3))
    info = np.array([[p[0, 0], p[0, 1], p[0, 2]], [p[0, 1], p[1, 1], p[1, 2]], [p[0, 2], p[1, 2], p[2, 2]]])
    return np.linalg.cholesky(info).T


def link_info(p1, p2, cov):
    """LinkInfo::Update (Mapper.h:174-188): pose2 in the frame of pose1 and the rotated covariance."""
    c, s = math.cos(p1[2]), math.sin(p1[2])
    dx, dy = p2[0] - p1[0], p2[1] - p1[1]
    diff = np.array([c * dx + s * dy, -s * dx + c * dy, p2[2] - p1[2]])
    t = -p1[2]
    R = np.array([[math.cos(t), -math.sin(t), 0.0], [math.sin(t), math.cos(t), 0.0], [0.0, 0.0, 1.0]])
    return diff, R @ np.asarray(cov).reshape(3, 3) @ R.T


@dataclass
class Options:
    """solvers/ceres_solver.cpp:96-186 (+ Ceres defaults where the reference sets nothing)."""
    max_num_iterations: int = 50
    function_tolerance: float = 1e-3
    gradient_tolerance: float = 1e-6
    parameter_tolerance: float = 1e-3
    min_relative_decrease: float = 1e-3
    initial_trust_region_radius: float = 1e4
    max_trust_region_radius: float = 1e8
    min_trust_region_radius: float = 1e-16
    min_lm_diagonal: float = 1e-6
    max_lm_diagonal: float = 1e32
    jacobi_scaling: bool = True
    use_nonmonotonic_steps: bool = True
    max_consecutive_nonmonotonic_steps: int = 3
    max_num_consecutive_invalid_steps: int = 3
    loss_function: str = "none"      # "none" | "huber" | "cauchy"  (solvers/ceres_solver.cpp:82-94)
    loss_scale: float = 0.7


@dataclass
class Summary:
    iterations: int = 0
    successful_steps: int = 0
    termination: str = ""
    usable: bool = True
    initial_cost: float = 0.0
    final_cost: float = 0.0
    trace: list = field(default_factory=list)   # (iteration, cost, accepted, radius)


class Problem:
    """Vectorised residual / Jacobian of PoseGraph2dErrorTerm over all edges."""

    def __init__(self, poses, edge_a, edge_b, z, U, fixed, loss="none", loss_scale=0.7):
        self.loss, self.loss_a = loss, loss_scale
        self.N = len(poses)
        self.ea = np.asarray(edge_a, dtype=np.int64)
        self.eb = np.asarray(edge_b, dtype=np.int64)
        self.z = np.asarray(z, dtype=np.float64)
        self.U = np.asarray(U, dtype=np.float64)            # (E,3,3) upper triangular
        # parameters in the problem: nodes touched by an edge, minus the constant anchor
        used = np.zeros(self.N, dtype=bool)
        used[self.ea] = True
        used[self.eb] = True
        if fixed is not None and fixed >= 0:
            used[fixed] = False
        self.free = np.nonzero(used)[0]
        self.col = -np.ones(self.N, dtype=np.int64)
        self.col[self.free] = np.arange(len(self.free))

    def _rho(self, sq):
        """ceres::HuberLoss / CauchyLoss Evaluate: rho(s), rho'(s)."""
        a, b = self.loss_a, self.loss_a ** 2
        tiny = np.finfo(np.float64).tiny
        if self.loss == "huber":
            r = np.sqrt(np.maximum(sq, tiny))
            big = sq > b
            return np.where(big, 2 * a * r - b, sq), np.where(big, np.maximum(tiny, a / r), 1.0)
        if self.loss == "cauchy":
            s = 1.0 + sq / b
            return b * np.log(s), np.maximum(tiny, 1.0 / s)
        return sq, np.ones_like(sq)

    def cost(self, x):
        r = self.raw_residuals(x).reshape(-1, 3)
        return 0.5 * float(self._rho((r * r).sum(axis=1))[0].sum())

    def residuals(self, x):
        """Residuals after Ceres' Corrector (rho'' <= 0 for both losses: scale by sqrt(rho'))."""
        r = self.raw_residuals(x).reshape(-1, 3)
        w = np.sqrt(self._rho((r * r).sum(axis=1))[1])
        return (r * w[:, None]).reshape(-1)

    def raw_residuals(self, x):
        pa, pb = x[self.ea], x[self.eb]
        c, s = np.cos(pa[:, 2]), np.sin(pa[:, 2])
        dx, dy = pb[:, 0] - pa[:, 0], pb[:, 1] - pa[:, 1]
        e = np.empty((len(self.ea), 3))
        e[:, 0] = c * dx + s * dy - self.z[:, 0]
        e[:, 1] = -s * dx + c * dy - self.z[:, 1]
        e[:, 2] = normalize_angle((pb[:, 2] - pa[:, 2]) - self.z[:, 2])
        return np.einsum("eij,ej->ei", self.U, e).reshape(-1)

    def jacobian(self, x):
        """Sparse (3E x 3F) Jacobian w.r.t. the free nodes' (x, y, theta)."""
        E = len(self.ea)
        pa, pb = x[self.ea], x[self.eb]
        c, s = np.cos(pa[:, 2]), np.sin(pa[:, 2])
        dx, dy = pb[:, 0] - pa[:, 0], pb[:, 1] - pa[:, 1]
        A = np.zeros((E, 3, 3))   # d e / d (xa, ya, tha)
        A[:, 0, 0] = -c; A[:, 0, 1] = -s; A[:, 0, 2] = -s * dx + c * dy
        A[:, 1, 0] = s; A[:, 1, 1] = -c; A[:, 1, 2] = -c * dx - s * dy
        A[:, 2, 2] = -1.0
        B = np.zeros((E, 3, 3))   # d e / d (xb, yb, thb)
        B[:, 0, 0] = c; B[:, 0, 1] = s
        B[:, 1, 0] = -s; B[:, 1, 1] = c
        B[:, 2, 2] = 1.0
        JA = np.einsum("eij,ejk->eik", self.U, A)
        JB = np.einsum("eij,ejk->eik", self.U, B)
        if self.loss != "none":
            r = self.raw_residuals(x).reshape(-1, 3)
            w = np.sqrt(self._rho((r * r).sum(axis=1))[1])
            JA = JA * w[:, None, None]
            JB = JB * w[:, None, None]
        rows = (3 * np.arange(E)[:, None, None] + np.arange(3)[None, :, None]) + np.zeros((1, 1, 3), dtype=np.int64)
        ca, cb = self.col[self.ea], self.col[self.eb]
        colsA = 3 * ca[:, None, None] + np.arange(3)[None, None, :] + np.zeros((1, 3, 1), dtype=np.int64)
        colsB = 3 * cb[:, None, None] + np.arange(3)[None, None, :] + np.zeros((1, 3, 1), dtype=np.int64)
        ma = np.broadcast_to((ca >= 0)[:, None, None], (E, 3, 3))
        mb = np.broadcast_to((cb >= 0)[:, None, None], (E, 3, 3))
        r = np.concatenate([rows[ma], rows[mb]])
        cc = np.concatenate([colsA[ma], colsB[mb]])
        v = np.concatenate([JA[ma], JB[mb]])
        return sp.csc_matrix((v, (r, cc)), shape=(3 * E, 3 * len(self.free)))

    def plus(self, x, delta):
        """x (+) delta on the free nodes: x, y plain add; theta through AngleLocalParameterization."""
        out = x.copy()
        d = delta.reshape(-1, 3)
        out[self.free, 0] += d[:, 0]
        out[self.free, 1] += d[:, 1]
        out[self.free, 2] = normalize_angle(out[self.free, 2] + d[:, 2])
        return out

    def params(self, x):
        return x[self.free].reshape(-1)


def solve(poses, edge_a, edge_b, z, cov=None, U=None, fixed=0, opts: Options | None = None, exact_solver="superlu"):
    """Ceres trust-region LM (see module docstring). Returns (optimised poses, Summary).
    poses (N,3); edges index into poses; cov (E,3,3) edge covariances (or U given directly)."""
    o = opts or Options()
    x = np.array(poses, dtype=np.float64)
    if U is None:
        U = np.stack([sqrt_information(c) for c in cov])
    pb = Problem(x, edge_a, edge_b, z, U, fixed, o.loss_function, o.loss_scale)
    sm = Summary()
    nfree = 3 * len(pb.free)
    if nfree == 0 or len(pb.ea) == 0:
        sm.termination = "CONVERGENCE (nothing to optimise)"
        return x, sm

    def evaluate(xx):
        return pb.cost(xx), pb.residuals(xx)

    def grad_and_jac(xx, r, scale):
        J = pb.jacobian(xx)
        g = J.T @ r                      # unscaled gradient (tangent space)
        if o.jacobi_scaling:
            if scale is None:
                scale = 1.0 / (1.0 + np.sqrt(np.asarray(J.multiply(J).sum(axis=0)).reshape(-1)))
            J = J @ sp.diags(scale)
        elif scale is None:
            scale = np.ones(nfree)
        xs = pb.params(xx)
        proj = pb.params(pb.plus(xx, -g))
        gmax = float(np.max(np.abs(xs - proj))) if len(xs) else 0.0
        return J, scale, gmax

    cost, r = evaluate(x)
    sm.initial_cost = cost
    J, scale, gmax = grad_and_jac(x, r, None)
    x_norm = float(np.linalg.norm(pb.params(x)))
    best_x, minimum_cost = x.copy(), cost

    radius, decrease_factor, reuse_diagonal = o.initial_trust_region_radius, 2.0, False
    diagonal = None
    # TrustRegionStepEvaluator state
    max_nonmono = o.max_consecutive_nonmonotonic_steps if o.use_nonmonotonic_steps else 0
    ev_min = ev_cur = ev_ref = ev_cand = cost
    acc_ref = acc_cand = 0.0
    n_nonmono = 0
    invalid_steps = 0
    it = 0
    step_successful = False
    sm.trace.append((0, cost, True, radius))
    # TrustRegionMinimizer::IterationZero: an already-converged start returns CONVERGENCE before any step is computed
    converged_at_start = gmax <= o.gradient_tolerance
    if converged_at_start:
        sm.termination = "CONVERGENCE (gradient tolerance)"
    while not converged_at_start:
        # FinalizeIterationAndCheckIfMinimizerCanContinue
        if it >= o.max_num_iterations:
            sm.termination = "NO_CONVERGENCE (max iterations)"
            break
        if step_successful and gmax <= o.gradient_tolerance:
            sm.termination = "CONVERGENCE (gradient tolerance)"
            break
        if radius <= o.min_trust_region_radius:
            sm.termination = "CONVERGENCE (min trust region radius)"
            break
        it += 1
        step_successful = False
        # LevenbergMarquardtStrategy::ComputeStep
        if not reuse_diagonal:
            diagonal = np.asarray(J.multiply(J).sum(axis=0)).reshape(-1)
            diagonal = np.minimum(np.maximum(diagonal, o.min_lm_diagonal), o.max_lm_diagonal)
        D2 = diagonal / radius
        H = (J.T @ J + sp.diags(D2)).tocsc()
        rhs = J.T @ r
        try:
            y = spla.splu(H).solve(rhs)
            ok = bool(np.all(np.isfinite(y)))
        except RuntimeError:
            ok = False
        reuse_diagonal = True
        valid = False
        if ok:
            step = -y
            mr = J @ step
            model_cost_change = -float(mr @ (r + mr / 2.0))
            valid = model_cost_change > 0.0
        if not valid:
            invalid_steps += 1
            if invalid_steps >= o.max_num_consecutive_invalid_steps:
                sm.termination = "FAILURE (too many invalid steps)"
                sm.usable = False
                break
            radius /= decrease_factor      # StepIsInvalid == StepRejected(0)
            decrease_factor *= 2.0
            sm.trace.append((it, cost, False, radius))
            continue
        invalid_steps = 0
        delta = step * scale
        cand = pb.plus(x, delta)
        cand_cost, cand_r = evaluate(cand)
        # ParameterToleranceReached
        step_norm = float(np.linalg.norm(pb.params(x) - pb.params(cand)))
        if step_norm <= o.parameter_tolerance * (x_norm + o.parameter_tolerance):
            sm.termination = "CONVERGENCE (parameter tolerance)"
            break
        # FunctionToleranceReached
        cost_change = cost - cand_cost
        if abs(cost_change) <= o.function_tolerance * cost:
            sm.termination = "CONVERGENCE (function tolerance)"
            break
        # IsStepSuccessful via TrustRegionStepEvaluator::StepQuality
        rel = (ev_cur - cand_cost) / model_cost_change
        hist = (ev_ref - cand_cost) / (acc_ref + model_cost_change)
        quality = max(rel, hist)
        if quality > o.min_relative_decrease:
            x, cost, r = cand, cand_cost, cand_r
            x_norm = float(np.linalg.norm(pb.params(x)))
            J, scale, gmax = grad_and_jac(x, r, scale)
            step_successful = True
            sm.successful_steps += 1
            radius = min(o.max_trust_region_radius, radius / max(1.0 / 3.0, 1.0 - (2.0 * quality - 1.0) ** 3))
            decrease_factor, reuse_diagonal = 2.0, False
            # StepAccepted
            ev_cur = cost
            acc_cand += model_cost_change
            acc_ref += model_cost_change
            if ev_cur < ev_min:
                ev_min = ev_cur; n_nonmono = 0; ev_cand = ev_cur; acc_cand = 0.0
            else:
                n_nonmono += 1
                if ev_cur > ev_cand:
                    ev_cand = ev_cur; acc_cand = 0.0
            if n_nonmono == max_nonmono:
                ev_ref = ev_cand; acc_ref = acc_cand
            if cost < minimum_cost:
                minimum_cost, best_x = cost, x.copy()
        else:
            radius /= decrease_factor
            decrease_factor *= 2.0
            reuse_diagonal = True
        sm.trace.append((it, cost, step_successful, radius))
    sm.iterations = it
    sm.final_cost = minimum_cost
    return (best_x if sm.usable else np.array(poses, dtype=np.float64)), sm
