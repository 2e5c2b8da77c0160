"""cfg5-shaped sweep on one GPU: Q queries x C candidate scans (all pairs). Prints pairs/s for the device-resident
run and the end-to-end call, plus the per-query winners."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from slam_toolbox_b200 import api, synth

Q = int(sys.argv[1]) if len(sys.argv) > 1 else 16
Cn = int(sys.argv[2]) if len(sys.argv) > 2 else 6250
world = synth.make_world(7)
rng = np.random.default_rng(5)
qtrue = synth.poses_near(world, synth.free_pose(world, rng)[:2], 2.0, Q, rng)
qr = synth.noisy(synth.raycast(world, qtrue), rng)
qp = qtrue + np.column_stack([rng.normal(0, 0.4, (Q, 2)), rng.normal(0, 0.06, Q)])
cposes = synth.poses_near(world, qtrue[0, :2], 3.0, Cn, rng)
cr = synth.noisy(synth.raycast(world, cposes, chunk=64), rng)
laser = api.LaserRangeFinder()
mapper = api.MapperParams(**{k: (bool(v) if k == "use_response_expansion" else v) for k, v in bench.LOOP_MAPPER.items()})
sm = api.ScanMatcher.Create(mapper, *bench.LOOP_GRID)
queries, cands = api.ScanBlock(qr, qp, laser), api.ScanBlock(cr, cposes, laser)
cs = np.arange(Cn + 1, dtype=np.int32)
t = time.perf_counter(); n = sm.batch_upload(queries, cands, cs, None, False); t_up = time.perf_counter() - t
print('info', sm.batch_info(), file=sys.stderr)
sm.batch_run(); k1 = sm.batch_kernel_ms()
t = time.perf_counter(); sm.batch_run(); k2 = sm.batch_kernel_ms(); t_run = time.perf_counter() - t
t = time.perf_counter(); resp, mean, cov = sm.batch_fetch(); t_fetch = time.perf_counter() - t
t = time.perf_counter(); r2 = sm.MatchScanBatch(queries, cands, cs, None, False, False); t_e2e = time.perf_counter() - t
assert np.array_equal(r2[0], resp)
best = resp.reshape(Q, Cn)
print(json.dumps({"pairs": n, "upload_s": t_up, "kernel_ms": k2, "pairs_per_s_kernel": n / (k2 * 1e-3), "fetch_s": t_fetch,
                  "e2e_s": t_e2e, "pairs_per_s_e2e": n / t_e2e, "best_response_per_query": np.round(best.max(axis=1), 3).tolist()}))
