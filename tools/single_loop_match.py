"""One loop-closure coarse match (chain of 10 scans, no penalty, no refinement) at the shipped 8 m search: the single-match path
(b200sm_match) against a one-pair batch on an 8-CTA cluster (b200sm_match_batch)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from slam_toolbox_b200 import api, synth
dim = float(sys.argv[1]) if len(sys.argv) > 1 else 8.0
sw = synth.make_loop_sweep(5, n_queries=1, n_chains=1, chain_len=10)
laser = api.LaserRangeFinder()
mapper = api.MapperParams(**{k: (bool(v) if k == "use_response_expansion" else v) for k, v in bench.LOOP_MAPPER.items()})
sm = api.ScanMatcher.Create(mapper, dim, 0.05, 0.03, 12.0)
q, c = api.ScanBlock(sw.query_ranges, sw.query_poses, laser), api.ScanBlock(sw.cand_ranges, sw.cand_poses, laser)
for _ in range(3):
    a = sm.MatchScan(q, c, False, False)
    b = sm.MatchScanBatch(q, c, sw.chain_start, None, False, False)
assert a[0] == b[0][0] and np.array_equal(a[1], b[1][0]) and np.array_equal(a[2], b[2][0])
sm.match_timing(reset=True)
t = time.perf_counter()
for _ in range(50):
    sm.MatchScan(q, c, False, False)
t1 = (time.perf_counter() - t) / 50
t = time.perf_counter()
for _ in range(50):
    sm.MatchScanBatch(q, c, sw.chain_start, None, False, False)
t2 = (time.perf_counter() - t) / 50
print("dim", dim, "single-match path %.3f ms" % (1e3 * t1), sm.match_timing(), "| one-pair batch %.3f ms" % (1e3 * t2), sm.batch_tile_info(), sm.batch_upload_timing(), "kernel ms", sm.batch_kernel_ms())
