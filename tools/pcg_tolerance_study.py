"""cfg4 at both noise levels for a range of PCG tolerances: solve time, iteration counts, and the largest pose difference to
the tightest solve -- how loose the inner solve may be before the LM trajectory (accept / reject sequence) changes."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from slam_toolbox_b200 import api, synth

for sig in ((0.05, 0.02), (0.03, 0.01)):
    g = synth.make_pose_graph(0, 10000, 40000, sigma_xy=sig[0], sigma_th=sig[1])
    ref = None
    for tol in (1e-12, 1e-10, 1e-9, 1e-8, 1e-7, 1e-6, 1e-5, 1e-4):
        s = api.ScanSolver(pcg_tolerance=tol)
        best = None
        for rep in range(2):
            s.Reset()
            for nid, p in zip(g["ids"], g["init"]):
                s.AddNode(int(nid), p)
            for a, b, z, c in zip(g["edge_a"], g["edge_b"], g["z"], g["cov"]):
                s.AddConstraint(int(a), int(b), z, c)
            s.Compute()
            best = s.summary.solve_ms
        x = s.GetCorrections()[1]
        if ref is None:
            ref = x
        d = x - ref
        d[:, 2] = synth.wrap(d[:, 2])
        print(json.dumps({"sigma": sig, "pcg_tolerance": tol, "ms": round(best, 2), "lm": int(s.summary.iterations), "ok": int(s.summary.successful_steps),
                          "pcg": int(s.summary.pcg_iterations), "final_cost": s.summary.final_cost,
                          "dxy": float(np.abs(d[:, :2]).max()), "dth": float(np.abs(d[:, 2]).max())}), flush=True)
        s.close()
