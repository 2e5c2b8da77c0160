"""Bring-up helper: cfg4-sized and smaller pose-graph solves against the oracle, with the solver's debug timers.
    B200PG_DEBUG=1 [B200PG_COARSE_MODES=3] python tools/gpu_quick_pg2.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from slam_toolbox_b200 import synth, api
from oracle import posegraph as PG

for n, e in ((300, 700), (2000, 6000), (10000, 40000)):
    g = synth.make_pose_graph(0, n, e, sigma_xy=0.03, sigma_th=0.01)
    xo, so = PG.solve(g["init"], g["edge_a"], g["edge_b"], g["z"], cov=g["cov"])
    for rep in range(2):
        s = api.ScanSolver()
        for i, p in zip(g["ids"], g["init"]): s.AddNode(int(i), p)
        for a, b, z, c in zip(g["edge_a"], g["edge_b"], g["z"], g["cov"]): s.AddConstraint(int(a), int(b), z, c)
        ok = s.Compute(); sm = s.summary
        d = s.GetCorrections()[1] - xo; d[:, 2] = synth.wrap(d[:, 2])
        print(f"n={n}: ok={ok} lm={sm.iterations} (oracle {so.iterations}) pcg={sm.pcg_iterations} solve_ms={sm.solve_ms:.2f} max|dxy|={np.abs(d[:,:2]).max():.2e} max|dth|={np.abs(d[:,2]).max():.2e}", flush=True)
        s.close()

# degenerate aggregates: tiny graphs (one aggregate, one or two free nodes), and isolated nodes inside an aggregate
for n, e in ((2, 1), (3, 3), (5, 7), (17, 30)):
    g = synth.make_pose_graph(2, max(n, 2), max(e, n - 1), sigma_xy=0.03, sigma_th=0.01, min_gap=1) if n > 3 else None
    s = api.ScanSolver()
    if g is None:
        poses = np.array([[0, 0, 0], [1.0, 0.1, 0.05], [2.1, 0.0, -0.02]])[:n]
        for i in range(n): s.AddNode(i, poses[i])
        cov = np.diag([0.01, 0.01, 0.001])
        for a, b in ([(0, 1)] if n == 2 else [(0, 1), (1, 2), (0, 2)]):
            s.AddConstraint(a, b, np.array([1.0, 0.0, 0.0]) * (b - a), cov)
        ok = s.Compute()
        print(f"tiny n={n}: ok={ok} lm={s.summary.iterations} pcg={s.summary.pcg_iterations} cost={s.summary.final_cost:.3e}", flush=True)
    else:
        xo, so = PG.solve(g["init"], g["edge_a"], g["edge_b"], g["z"], cov=g["cov"])
        for i, p in zip(g["ids"], g["init"]): s.AddNode(int(i), p)
        for k in range(3): s.AddNode(1000 + k, np.zeros(3))          # isolated nodes: never parameters
        for a, b, z, c in zip(g["edge_a"], g["edge_b"], g["z"], g["cov"]): s.AddConstraint(int(a), int(b), z, c)
        ok = s.Compute()
        ids, xg = s.GetCorrections()
        d = xg[:len(xo)] - xo; d[:, 2] = synth.wrap(d[:, 2])
        print(f"small n={n}: ok={ok} lm={s.summary.iterations} (oracle {so.iterations}) pcg={s.summary.pcg_iterations} max|dxy|={np.abs(d[:,:2]).max():.2e}", flush=True)
    s.close()
