"""Ad-hoc GPU bring-up check of the pose-graph solver against the restated-Ceres oracle."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from slam_toolbox_b200 import synth, api
from oracle import posegraph as PG

def run(n, e, seed=0, **kw):
    g = synth.make_pose_graph(seed, n, e, sigma_xy=0.03, sigma_th=0.01)
    t = time.time(); xo, so = PG.solve(g["init"], g["edge_a"], g["edge_b"], g["z"], cov=g["cov"]); to = time.time() - t
    s = api.ScanSolver(**kw)
    for i, p in zip(g["ids"], g["init"]): s.AddNode(int(i), p)
    for a, b, z, c in zip(g["edge_a"], g["edge_b"], g["z"], g["cov"]): s.AddConstraint(int(a), int(b), z, c)
    t = time.time(); ok = s.Compute(); tg = time.time() - t
    ids, xg = s.GetCorrections()
    sm = s.summary
    d = xg - xo; d[:, 2] = synth.wrap(d[:, 2] + np.pi) - np.pi
    print(f"n={n} e={len(g['edge_a'])}: oracle {so.termination} it={so.iterations} acc={so.successful_steps} cost={so.final_cost:.6g} ({to:.2f}s) | "
          f"gpu ok={ok} term={sm.termination} it={sm.iterations} acc={sm.successful_steps} pcg={sm.pcg_iterations} cost={sm.final_cost:.6g} "
          f"solve_ms={sm.solve_ms:.2f} wall={tg*1e3:.1f}ms launches={sm.kernel_launches} | max|dx|={np.abs(d[:,:2]).max():.3e} max|dth|={np.abs(d[:,2]).max():.3e}", flush=True)
    # second solve from the solved state (warm): should converge immediately
    ok = s.Compute(); sm = s.summary
    print(f"   re-solve: term={sm.termination} it={sm.iterations} pcg={sm.pcg_iterations} solve_ms={sm.solve_ms:.2f}")

run(200, 500)
run(2000, 6000)
run(10000, 40000)
run(10000, 40000, pcg_tolerance=1e-12)
