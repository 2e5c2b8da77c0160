"""Realistic graph shape (4,265 vertices / ~5,210 edges in the reference's own recorded run): solve against the oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from slam_toolbox_b200 import synth, api
from oracle import posegraph as PG

for seed, n, e in ((5, 4265, 5210), (6, 1000, 1150)):
    g = synth.make_pose_graph(seed, n, e, sigma_xy=0.03, sigma_th=0.01)
    xo, so = PG.solve(g["init"], g["edge_a"], g["edge_b"], g["z"], cov=g["cov"])
    s = api.ScanSolver()
    for i, p in zip(g["ids"], g["init"]): s.AddNode(int(i), p)
    for a, b, z, c in zip(g["edge_a"], g["edge_b"], g["z"], g["cov"]): s.AddConstraint(int(a), int(b), z, c)
    ok = s.Compute(); sm = s.summary
    d = s.GetCorrections()[1] - xo; d[:, 2] = synth.wrap(d[:, 2])
    print(f"n={n} e={len(g['edge_a'])}: ok={ok} lm={sm.iterations} (oracle {so.iterations}) pcg={sm.pcg_iterations} solve_ms={sm.solve_ms:.2f} "
          f"cost={sm.final_cost:.6g} (oracle {so.final_cost:.6g}) max|dxy|={np.abs(d[:,:2]).max():.2e} max|dth|={np.abs(d[:,2]).max():.2e}", flush=True)
