"""Small driver for ncu: one cfg4 pose-graph solve (10k nodes / 40k edges)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from slam_toolbox_b200 import api, synth

g = synth.make_pose_graph(0, 10000, 40000, sigma_xy=0.03, sigma_th=0.01)
s = api.ScanSolver()
for nid, p in zip(g["ids"], g["init"]):
    s.AddNode(int(nid), p)
for a, b, z, c in zip(g["edge_a"], g["edge_b"], g["z"], g["cov"]):
    s.AddConstraint(int(a), int(b), z, c)
print("ok", s.Compute(), "ms", s.summary.solve_ms, "pcg", s.summary.pcg_iterations)
