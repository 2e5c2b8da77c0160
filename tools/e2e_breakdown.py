"""Host-side breakdown of the end-to-end batched sweep (cfg2): upload / run / fetch wall times."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from slam_toolbox_b200 import api
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
kern = int(sys.argv[2]) if len(sys.argv) > 2 else 0
qr, qp, cr, cp, cs = bench.make_inputs(0, n, 1, 1)
laser = api.LaserRangeFinder()
mapper = api.MapperParams(**{k: (bool(v) if k == "use_response_expansion" else v) for k, v in bench.LOOP_MAPPER.items()})
sm = api.ScanMatcher.Create(mapper, *bench.LOOP_GRID)
sm.set_option("sweep_kernel", kern)
pts = torch.empty((cr.shape[0], cr.shape[1], 2), dtype=torch.float64).pin_memory()
pts.numpy()[...] = api.point_readings(cr, cp, laser)
c, q = api.ScanBlock(cr, cp, laser, points=pts.numpy()), api.ScanBlock(qr, qp, laser)
for _ in range(3):
    sm.MatchScanBatch(q, c, cs, None, False, False)
T = {"upload": [], "run+sync": [], "fetch": [], "fused": []}
for _ in range(10):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); sm.batch_upload(q, c, cs, None, False); t1 = time.perf_counter()
    sm.batch_run(); k = sm.batch_kernel_ms(); t2 = time.perf_counter()
    sm.batch_fetch(); t3 = time.perf_counter()
    sm.MatchScanBatch(q, c, cs, None, False, False); t4 = time.perf_counter()
    T["upload"].append(t1 - t0); T["run+sync"].append(t2 - t1); T["fetch"].append(t3 - t2); T["fused"].append(t4 - t3)
print({k: round(1e3 * float(np.median(v)), 3) for k, v in T.items()}, "kernel_ms", round(k, 3), sm.batch_info()["kernel"], sm.batch_upload_timing())
