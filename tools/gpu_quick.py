"""Ad-hoc GPU bring-up check (not a pytest file): CUDA path vs C-port oracle on a few cases."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from slam_toolbox_b200 import synth
import helpers as H

def cmp(name, a, b):
    ok = all(np.array_equal(np.asarray(x), np.asarray(y)) for x, y in zip(a, b))
    print(f"{name}: {'EXACT' if ok else 'MISMATCH'}", flush=True)
    if not ok:
        for x, y in zip(a, b):
            print("   ", np.asarray(x).ravel()[:6], np.asarray(y).ravel()[:6])
    return ok

allok = True
for grid_name, mapper, grid in (("SEQ", H.MAPPER_SEQ, H.GRID_SEQ), ("SEQ_YAML", H.MAPPER_SEQ, H.GRID_SEQ_YAML), ("LOOP", H.MAPPER_LOOP, H.GRID_LOOP)):
    case = synth.make_sequential_case(1, inf_frac=0.03)
    pm, gm = H.port_matcher(mapper, grid), H.gpu_matcher(mapper, grid)
    pb, pq = H.port_scans(case["base_ranges"], case["base_poses"]), H.port_scans(case["query_ranges"], case["query_pose"])[0]
    gb, gq = H.gpu_block(case["base_ranges"], case["base_poses"]), H.gpu_block(case["query_ranges"], case["query_pose"])
    for pen, ref in ((True, True), (False, False)):
        t = time.time(); r1 = pm.match(pq, pb, pen, ref); t1 = time.time() - t
        t = time.time(); r2 = gm.MatchScan(gq, gb, pen, ref); t2 = time.time() - t
        t = time.time(); r2 = gm.MatchScan(gq, gb, pen, ref); t3 = time.time() - t
        allok &= cmp(f"{grid_name} match pen={pen} refine={ref} (cpu {t1*1e3:.1f} ms, gpu first {t2*1e3:.1f} ms, again {t3*1e3:.2f} ms) resp={r1[0]:.4f}", r1, r2)
    allok &= cmp(f"{grid_name} grid bytes", [pm.grid()["data"]], [gm.GetCorrelationGrid()["data"]])
    off, res = H.coarse_search(grid)
    pm.raster(pq, pb); gm.raster(gq, gb)
    c1 = pm.correlate(pq, case["query_pose"], off, res, mapper["coarse_search_angle_offset"], mapper["coarse_angle_resolution"], True, False)
    c2 = gm.CorrelateScan(gq, case["query_pose"], off, res, mapper["coarse_search_angle_offset"], mapper["coarse_angle_resolution"], True, False)
    allok &= cmp(f"{grid_name} coarse volume {c1[3].shape}", c1, c2)

# batch sweep
for chain_len, nch in ((1, 64), (10, 24)):
    sw = synth.make_loop_sweep(3, n_queries=2, n_chains=nch, chain_len=chain_len, inf_frac=0.02)
    pm, gm = H.port_matcher(H.MAPPER_LOOP, H.GRID_LOOP), H.gpu_matcher(H.MAPPER_LOOP, H.GRID_LOOP)
    pc, pq = H.port_scans(sw.cand_ranges, sw.cand_poses), H.port_scans(sw.query_ranges, sw.query_poses)
    gc, gq = H.gpu_block(sw.cand_ranges, sw.cand_poses), H.gpu_block(sw.query_ranges, sw.query_poses)
    for refine in (False, True):
        t = time.time()
        exp = [pm.match(pq[q], pc[sw.chain_start[c]:sw.chain_start[c + 1]], False, refine) for q in range(2) for c in range(nch)]
        t1 = time.time() - t
        t = time.time(); resp, mean, cov = gm.MatchScanBatch(gq, gc, sw.chain_start, None, False, refine); t2 = time.time() - t
        allok &= cmp(f"batch chain={chain_len} refine={refine} pairs={len(exp)} cpu {t1:.2f}s gpu {t2*1e3:.1f} ms",
                     (np.array([e[0] for e in exp]), np.array([e[1] for e in exp]), np.array([e[2] for e in exp])), (resp, mean, cov))
    print("  responses:", np.round(resp[:8], 3), "kernel ms", gm.batch_kernel_ms(), "launches", gm.launch_count())
print("ALL OK" if allok else "FAILURES")
