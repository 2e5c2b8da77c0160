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
