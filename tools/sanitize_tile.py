"""compute-sanitizer target: a small sweep through the tiled cluster kernel (4-CTA clusters and one CTA per pair, two geometries)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import helpers as H
from slam_toolbox_b200 import synth
sw = synth.make_loop_sweep(31, n_queries=2, n_chains=5, chain_len=2, inf_frac=0.02)
mapper = dict(H.MAPPER_LOOP, use_response_expansion=0)
for grid in (H.GRID_LOOP, (8.0, 0.05, 0.03, 20.0), (4.0, 0.05, 0.03, 6.0)):
    gm = H.gpu_matcher(mapper, grid)
    gc, gq = H.gpu_block(sw.cand_ranges, sw.cand_poses), H.gpu_block(sw.query_ranges, sw.query_poses)
    ref = None
    for cl in (1, 4):
        gm.set_option("sweep_kernel", 2); gm.set_option("sweep_cluster", cl)
        r = gm.MatchScanBatch(gq, gc, sw.chain_start, None, False, False)
        if ref is None:
            ref = r
        assert all(np.array_equal(a, b) for a, b in zip(r, ref))
    print(grid, gm.batch_tile_info(), float(r[0].max()))
