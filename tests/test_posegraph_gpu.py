"""Parity of the CUDA pose-graph solver with the restated-Ceres oracle (tolerance from BASELINE.json's
north_star: 1e-4 m / 1e-5 rad) and the ScanSolver API semantics of solvers/ceres_solver.cpp."""
import numpy as np
import pytest

from oracle import posegraph as PG
from slam_toolbox_b200 import api, synth

pytestmark = pytest.mark.gpu
TOL_XY, TOL_TH = 1e-4, 1e-5


def build(g, **opts):
    s = api.ScanSolver(**opts)
    for nid, p in zip(g["ids"], g["init"]):
        s.AddNode(int(nid), p)
    for a, b, z, c in zip(g["edge_a"], g["edge_b"], g["z"], g["cov"]):
        assert s.AddConstraint(int(a), int(b), z, c)
    return s


def diff(x, y):
    d = x - y
    d[:, 2] = synth.wrap(d[:, 2])
    return np.abs(d[:, :2]).max(), np.abs(d[:, 2]).max()


@pytest.mark.parametrize("n,e,seed", [(60, 120, 0), (500, 1400, 1), (3000, 9000, 2)])
def test_solve_matches_oracle_at_reference_tolerances(n, e, seed):
    g = synth.make_pose_graph(seed, n, e, sigma_xy=0.03, sigma_th=0.01)
    xo, so = PG.solve(g["init"], g["edge_a"], g["edge_b"], g["z"], cov=g["cov"])
    s = build(g)
    assert s.Compute()
    ids, xg = s.GetCorrections()
    assert np.array_equal(ids, g["ids"])
    dxy, dth = diff(xg, xo)
    assert dxy < TOL_XY and dth < TOL_TH, (dxy, dth)
    assert s.summary.iterations == so.iterations and s.summary.successful_steps == so.successful_steps
    assert abs(s.summary.final_cost - so.final_cost) <= 1e-8 * so.final_cost + 1e-18
    assert np.array_equal(xg[0], g["init"][0])   # first node is the constant anchor


def test_solve_matches_oracle_at_tight_tolerances():
    g = synth.make_pose_graph(5, 400, 1100, sigma_xy=0.03, sigma_th=0.01)
    kw = dict(function_tolerance=1e-14, parameter_tolerance=1e-13, gradient_tolerance=1e-13, max_num_iterations=100)
    xo, so = PG.solve(g["init"], g["edge_a"], g["edge_b"], g["z"], cov=g["cov"], opts=PG.Options(**kw))
    s = build(g, pcg_tolerance=1e-13, **kw)
    assert s.Compute()
    dxy, dth = diff(s.GetCorrections()[1], xo)
    assert dxy < 1e-6 and dth < 1e-6, (dxy, dth)


@pytest.mark.parametrize("sigma", [(0.03, 0.01), (0.05, 0.02)])
def test_cfg4_full_size(sigma):
    """cfg4 at both noise levels: SURVEY.md 8(d) fixes 0.05 m / 0.02 rad; 0.03 / 0.01 is the lower-noise variant the bench also
    reports.  P1 of the parity protocol: same LM iteration / accepted-step counts and final cost as the exact-solve oracle at the
    reference's tolerances, poses within 1e-4 m / 1e-5 rad."""
    g = synth.make_pose_graph(0, 10000, 40000, sigma_xy=sigma[0], sigma_th=sigma[1])
    xo, so = PG.solve(g["init"], g["edge_a"], g["edge_b"], g["z"], cov=g["cov"])
    s = build(g)
    assert s.Compute()
    dxy, dth = diff(s.GetCorrections()[1], xo)
    assert s.summary.iterations == so.iterations and s.summary.successful_steps == so.successful_steps, (s.summary.iterations, so.iterations)
    assert abs(s.summary.final_cost - so.final_cost) <= 1e-8 * so.final_cost
    assert dxy < TOL_XY and dth < TOL_TH, (dxy, dth)
    if sigma[0] < 0.04:
        # size-independent property: cost near (3E - 3(N-1))/2 for unit-variance whitened residuals (the global minimum)
        assert 0.8 < s.summary.final_cost / (0.5 * (3 * 40000 - 3 * 9999)) < 1.2
    assert s.Compute() and s.summary.iterations <= 2     # re-solve converges at once
    assert s.summary.uploaded_edges == 0                 # ... and uploads no constraint again


def test_incremental_solves_upload_only_new_constraints():
    """The mapper's traffic (Mapper.cpp:2012-2030: Compute after every loop closure on a graph that only grew): the second Compute
    uploads the appended constraints only and gives exactly what a fresh solver gives from the same state."""
    g = synth.make_pose_graph(8, 1500, 4200, sigma_xy=0.03, sigma_th=0.01)
    n1 = 1000
    first = (g["edge_a"] < n1) & (g["edge_b"] < n1)
    s = api.ScanSolver()
    for nid, p in zip(g["ids"][:n1], g["init"][:n1]):
        s.AddNode(int(nid), p)
    for a, b, z, c in zip(g["edge_a"][first], g["edge_b"][first], g["z"][first], g["cov"][first]):
        assert s.AddConstraint(int(a), int(b), z, c)
    assert s.Compute() and s.summary.uploaded_edges == int(first.sum())
    mid = s.GetCorrections()[1].copy()
    for nid, p in zip(g["ids"][n1:], g["init"][n1:]):
        s.AddNode(int(nid), p)
    for a, b, z, c in zip(g["edge_a"][~first], g["edge_b"][~first], g["z"][~first], g["cov"][~first]):
        assert s.AddConstraint(int(a), int(b), z, c)
    assert s.Compute() and s.summary.uploaded_edges == int((~first).sum())
    # the same state in a fresh solver (constraints in the same order)
    f = api.ScanSolver()
    for nid, p in zip(g["ids"], np.vstack([mid, g["init"][n1:]])):
        f.AddNode(int(nid), p)
    order = np.concatenate([np.nonzero(first)[0], np.nonzero(~first)[0]])
    for k in order:
        assert f.AddConstraint(int(g["edge_a"][k]), int(g["edge_b"][k]), g["z"][k], g["cov"][k])
    assert f.Compute()
    assert np.array_equal(s.GetCorrections()[1], f.GetCorrections()[1])
    assert (s.summary.iterations, s.summary.pcg_iterations) == (f.summary.iterations, f.summary.pcg_iterations)
    # a removal invalidates the device copy: everything is uploaded again, results still equal a fresh solver's
    s.RemoveConstraint(int(g["edge_a"][order[-1]]), int(g["edge_b"][order[-1]]))
    assert s.Compute() and s.summary.uploaded_edges == len(order) - 1


def test_api_semantics():
    g = synth.make_pose_graph(6, 50, 90, sigma_xy=0.03, sigma_th=0.01)
    s = build(g)
    assert s.num_nodes() == 50
    # unknown / identical nodes are refused, state unchanged (ceres_solver.cpp:351-358)
    ne = s.num_edges()
    assert not s.AddConstraint(0, 999, g["z"][0], g["cov"][0]) and not s.AddConstraint(3, 3, g["z"][0], g["cov"][0])
    assert s.num_edges() == ne
    assert len(s.GetCorrections()[0]) == 0          # nothing before Compute
    assert s.Compute()
    assert len(s.GetCorrections()[0]) == 50
    s.Clear()
    assert len(s.GetCorrections()[0]) == 0 and s.num_nodes() == 50
    # ModifyNode adds the stored yaw (ceres_solver.cpp:457-459)
    before = s.get_node(7)
    s.ModifyNode(7, [1.0, 2.0, 0.25])
    after = s.get_node(7)
    assert after[0] == 1.0 and after[1] == 2.0 and after[2] == 0.25 + before[2]
    assert s.GetNodeOrientation(7) == after[2] and s.GetNodeOrientation(12345) is None
    # RemoveConstraint in either orientation, RemoveNode drops its edges
    a, b = int(g["edge_a"][5]), int(g["edge_b"][5])
    assert s.RemoveConstraint(b, a) and not s.RemoveConstraint(b, a)
    ne = s.num_edges()
    assert s.RemoveNode(20) and not s.RemoveNode(20)
    assert s.num_nodes() == 49 and s.num_edges() < ne
    assert s.Compute() and len(s.GetCorrections()[0]) == 49
    # an isolated node is not part of the problem and comes back unchanged
    s.AddNode(777, [9.0, 9.0, 0.5])
    assert s.Compute()
    ids, poses = s.GetCorrections()
    assert np.array_equal(poses[list(ids).index(777)], [9.0, 9.0, 0.5])
    s.Reset()
    assert s.num_nodes() == 0 and s.num_edges() == 0 and len(s.GetCorrections()[0]) == 0


def test_removed_graph_equals_freshly_built_graph():
    g = synth.make_pose_graph(8, 120, 300, sigma_xy=0.03, sigma_th=0.01)
    keep = np.ones(len(g["edge_a"]), dtype=bool)
    keep[150:170] = False
    s1 = build(g)
    for k in np.nonzero(~keep)[0]:
        assert s1.RemoveConstraint(int(g["edge_a"][k]), int(g["edge_b"][k]))
    g2 = dict(g, edge_a=g["edge_a"][keep], edge_b=g["edge_b"][keep], z=g["z"][keep], cov=g["cov"][keep])
    s2 = build(g2)
    assert s1.Compute() and s2.Compute()
    assert np.array_equal(s1.GetCorrections()[1], s2.GetCorrections()[1])   # deterministic, no atomics


def test_smem_and_global_pcg_kernels_agree(monkeypatch):
    """The shared-memory-resident PCG kernel against the global-memory one (forced through the env switch)."""
    g = synth.make_pose_graph(11, 1500, 4500, sigma_xy=0.03, sigma_th=0.01)
    s1 = build(g)
    assert s1.Compute()
    monkeypatch.setenv("B200PG_FORCE_GLOBAL_PCG", "1")
    s2 = build(g)
    assert s2.Compute()
    assert s1.summary.iterations == s2.summary.iterations
    dxy, dth = diff(s1.GetCorrections()[1], s2.GetCorrections()[1])
    assert dxy < 1e-7 and dth < 1e-8, (dxy, dth)


def test_two_level_and_jacobi_preconditioners_agree(monkeypatch):
    """Two-level (rigid-mode aggregation) PCG against block-Jacobi PCG: same LM trajectory, same poses to PCG accuracy,
    far fewer CG iterations."""
    g = synth.make_pose_graph(12, 4000, 14000, sigma_xy=0.03, sigma_th=0.01)
    s1 = build(g)
    assert s1.Compute()
    monkeypatch.setenv("B200PG_PRECOND", "jacobi")
    s2 = build(g)
    assert s2.Compute()
    assert s1.summary.iterations == s2.summary.iterations and s1.summary.successful_steps == s2.summary.successful_steps
    dxy, dth = diff(s1.GetCorrections()[1], s2.GetCorrections()[1])
    assert dxy < 1e-7 and dth < 1e-8, (dxy, dth)
    assert s1.summary.pcg_iterations < 0.6 * s2.summary.pcg_iterations, (s1.summary.pcg_iterations, s2.summary.pcg_iterations)


def test_realistic_sparse_graph_shape():
    """The reference's own recorded mapping run has 4,265 vertices and ~5,210 edges (test/constraints_on_graph.dat):
    a long chain with few loop closures, far sparser than cfg4."""
    g = synth.make_pose_graph(5, 4265, 5210, sigma_xy=0.03, sigma_th=0.01)
    xo, so = PG.solve(g["init"], g["edge_a"], g["edge_b"], g["z"], cov=g["cov"])
    s = build(g)
    assert s.Compute()
    assert s.summary.iterations == so.iterations and abs(s.summary.final_cost - so.final_cost) <= 1e-9 * so.final_cost
    dxy, dth = diff(s.GetCorrections()[1], xo)
    assert dxy < TOL_XY and dth < TOL_TH, (dxy, dth)


def test_tiny_and_degenerate_graphs_match_oracle():
    """Aggregates the coarse space must survive: a single free node (the linear modes would repeat the rigid ones), three
    nodes, and isolated (never optimised) nodes sharing an aggregate with free ones."""
    cov = np.diag([0.01, 0.01, 0.001])
    for n, edges in ((2, [(0, 1)]), (3, [(0, 1), (1, 2), (0, 2)])):
        poses = np.array([[0, 0, 0], [1.05, 0.1, 0.05], [2.1, -0.05, -0.02]])[:n]
        ea, eb = np.array([a for a, _ in edges]), np.array([b for _, b in edges])
        z = np.array([[float(b - a), 0.0, 0.0] for a, b in edges])
        xo, so = PG.solve(poses, ea, eb, z, cov=np.repeat(cov[None], len(edges), axis=0))
        s = api.ScanSolver()
        for i in range(n):
            s.AddNode(i, poses[i])
        for (a, b), zz in zip(edges, z):
            assert s.AddConstraint(a, b, zz, cov)
        assert s.Compute()
        dxy, dth = diff(s.GetCorrections()[1], xo)
        assert dxy < TOL_XY and dth < TOL_TH, (n, dxy, dth)
    g = synth.make_pose_graph(21, 40, 70, sigma_xy=0.03, sigma_th=0.01, min_gap=3)
    xo, so = PG.solve(g["init"], g["edge_a"], g["edge_b"], g["z"], cov=g["cov"])
    s = api.ScanSolver()
    for k, (nid, p) in enumerate(zip(g["ids"], g["init"])):
        s.AddNode(int(nid), p)
        if k % 3 == 0:
            s.AddNode(10000 + k, np.array([5.0, 5.0, 0.3]))       # isolated: stays where it is
    for a, b, zz, c in zip(g["edge_a"], g["edge_b"], g["z"], g["cov"]):
        assert s.AddConstraint(int(a), int(b), zz, c)
    assert s.Compute() and s.summary.iterations == so.iterations
    ids, x = s.GetCorrections()
    keep = ids < 10000
    dxy, dth = diff(x[keep], xo)
    assert dxy < TOL_XY and dth < TOL_TH, (dxy, dth)
    assert np.array_equal(x[~keep], np.tile([5.0, 5.0, 0.3], ((~keep).sum(), 1)))


def test_linear_coarse_modes_against_rigid_only(monkeypatch):
    """6 coarse modes per aggregate (rigid + piecewise-linear deformation, the default) against the rigid-only coarse space:
    same LM trajectory, same poses to PCG accuracy, fewer CG iterations."""
    g = synth.make_pose_graph(13, 6000, 22000, sigma_xy=0.03, sigma_th=0.01)
    s1 = build(g)
    assert s1.Compute()
    monkeypatch.setenv("B200PG_COARSE_MODES", "3")
    s2 = build(g)
    assert s2.Compute()
    assert s1.summary.iterations == s2.summary.iterations and s1.summary.successful_steps == s2.summary.successful_steps
    dxy, dth = diff(s1.GetCorrections()[1], s2.GetCorrections()[1])
    assert dxy < 1e-7 and dth < 1e-8, (dxy, dth)
    assert s1.summary.pcg_iterations < 0.85 * s2.summary.pcg_iterations, (s1.summary.pcg_iterations, s2.summary.pcg_iterations)


@pytest.mark.parametrize("loss,code", [("huber", 1), ("cauchy", 2)])
def test_robust_losses_match_oracle_and_reject_outliers(loss, code):
    """ceres_loss_function = HuberLoss / CauchyLoss (scale 0.7, solvers/ceres_solver.cpp:82-94) with gross outliers
    injected into some loop-closure edges."""
    g = synth.make_pose_graph(21, 2500, 9000, sigma_xy=0.03, sigma_th=0.01)
    rng = np.random.default_rng(0)
    z = g["z"].copy()
    loops = np.arange(2499, len(z))
    bad = rng.choice(loops, size=max(4, len(loops) // 50), replace=False)
    z[bad, :2] += rng.normal(0, 2.0, (len(bad), 2))
    g = dict(g, z=z)
    xo, so = PG.solve(g["init"], g["edge_a"], g["edge_b"], g["z"], cov=g["cov"], opts=PG.Options(loss_function=loss))
    s = build(g, loss_function=code, loss_scale=0.7)
    assert s.Compute()
    dxy, dth = diff(s.GetCorrections()[1], xo)
    assert dxy < TOL_XY and dth < TOL_TH, (dxy, dth)
    assert s.summary.iterations == so.iterations
    assert abs(s.summary.final_cost - so.final_cost) <= 1e-8 * so.final_cost
    # the robust solution is closer to the truth than the squared-loss one
    s0 = build(g)
    assert s0.Compute()
    e_rob = np.abs(s.GetCorrections()[1][:, :2] - g["truth"][:, :2]).max()
    e_sq = np.abs(s0.GetCorrections()[1][:, :2] - g["truth"][:, :2]).max()
    assert e_rob < e_sq
