"""Parity of the CUDA scan matcher (through the C ABI) with the oracle: bit-exact integer correlation
volumes, arg-max indices, grid bytes, and -- because the FP64 epilogue keeps the reference's operation
order -- bit-exact response / mean / covariance.  Tolerances: none (np.array_equal)."""
import hashlib
import os

import numpy as np
import pytest

import helpers as H
from slam_toolbox_b200 import api, synth

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "matcher_golden.npz")
CONFIG = {"seq_k03": (H.MAPPER_SEQ, H.GRID_SEQ), "seq_yaml_inf": (H.MAPPER_SEQ, H.GRID_SEQ_YAML),
          "loop_chain5": (H.MAPPER_LOOP, H.GRID_LOOP), "loop_refine": (H.MAPPER_LOOP, H.GRID_LOOP),
          "small": (H.MAPPER_LOOP, H.GRID_SMALL)}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def same(a, b):
    return a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])


@pytest.mark.parametrize("name", sorted(CONFIG))
def test_golden_fixtures_from_the_reference(name):
    z = np.load(GOLDEN)
    mapper, grid = CONFIG[name]
    gm = H.gpu_matcher(mapper, grid)
    base = H.gpu_block(z[f"{name}/base_ranges"], z[f"{name}/base_poses"])
    q = H.gpu_block(z[f"{name}/query_ranges"], z[f"{name}/query_pose"])
    pen, refine = (bool(v) for v in z[f"{name}/flags"])
    resp, mean, cov = gm.MatchScan(q, base, pen, refine)
    assert resp == z[f"{name}/response"][0]
    assert np.array_equal(mean, z[f"{name}/mean"]) and np.array_equal(cov, z[f"{name}/cov"])
    gm.raster(q, base)
    assert sha(gm.GetCorrelationGrid()["data"]) == z[f"{name}/grid_sha"][0]
    so, sr = H.coarse_search(grid)
    _, _, _, vol = gm.CorrelateScan(q, z[f"{name}/query_pose"], so, sr, mapper["coarse_search_angle_offset"],
                                    mapper["coarse_angle_resolution"], False, False)
    assert sha(vol) == z[f"{name}/volume_sha"][0]
    assert int(vol.argmax()) == z[f"{name}/volume_argmax"][0] and int(vol.max()) == z[f"{name}/volume_argmax"][1]


@pytest.mark.parametrize("cfg", ["seq", "seq_yaml", "loop"])
@pytest.mark.parametrize("seed", [0, 1])
def test_single_match_vs_oracle(cfg, seed):
    mapper, grid = {"seq": (H.MAPPER_SEQ, H.GRID_SEQ), "seq_yaml": (H.MAPPER_SEQ, H.GRID_SEQ_YAML),
                    "loop": (H.MAPPER_LOOP, H.GRID_LOOP)}[cfg]
    case = synth.make_sequential_case(200 + seed, buffer_len=6, inf_frac=0.03, nan_frac=0.01 * seed)
    pm, gm = H.port_matcher(mapper, grid), H.gpu_matcher(mapper, grid)
    pb, pq = H.port_scans(case["base_ranges"], case["base_poses"]), H.port_scans(case["query_ranges"], case["query_pose"])[0]
    gb, gq = H.gpu_block(case["base_ranges"], case["base_poses"]), H.gpu_block(case["query_ranges"], case["query_pose"])
    for pen, refine in ((True, True), (False, False), (True, False), (False, True)):
        assert same(pm.match(pq, pb, pen, refine), gm.MatchScan(gq, gb, pen, refine))
    assert np.array_equal(pm.grid()["data"], gm.GetCorrelationGrid()["data"])
    so, sr = H.coarse_search(grid)
    pm.raster(pq, pb); gm.raster(gq, gb)
    for pen in (False, True):
        a = pm.correlate(pq, case["query_pose"], so, sr, mapper["coarse_search_angle_offset"], mapper["coarse_angle_resolution"], pen, False)
        b = gm.CorrelateScan(gq, case["query_pose"], so, sr, mapper["coarse_search_angle_offset"], mapper["coarse_angle_resolution"], pen, False)
        assert same(a, b) and np.array_equal(a[3], b[3])


def test_edge_cases_vs_oracle():
    mapper, grid = H.MAPPER_LOOP, H.GRID_SMALL
    pm, gm = H.port_matcher(mapper, grid), H.gpu_matcher(mapper, grid)
    case = synth.make_sequential_case(7, buffer_len=2)
    pb, gb = H.port_scans(case["base_ranges"], case["base_poses"]), H.gpu_block(case["base_ranges"], case["base_poses"])
    pq, gq = H.port_scans(case["query_ranges"], case["query_pose"])[0], H.gpu_block(case["query_ranges"], case["query_pose"])
    # no base scans: zero response, all poses tie, response expansion runs (use_response_expansion = 1)
    assert same(pm.match(pq, [], True, True), gm.MatchScan(gq, None, True, True))
    # every query reading invalid
    bad = np.full_like(case["query_ranges"], np.inf)
    assert same(pm.match(H.port_scans(bad, case["query_pose"])[0], pb, False, False), gm.MatchScan(H.gpu_block(bad, case["query_pose"]), gb, False, False))
    # query far from the base scans
    far = case["query_pose"] + np.array([500.0, -300.0, 1.0])
    assert same(pm.match(H.port_scans(case["query_ranges"], far)[0], pb, True, False), gm.MatchScan(H.gpu_block(case["query_ranges"], far), gb, True, False))
    # a scan with zero readings (Mapper.cpp:547-557)
    empty = api.ScanBlock(np.zeros((1, 0)), case["query_pose"][None, :], api.LaserRangeFinder())
    r, m, c = gm.MatchScan(empty, gb, True, True)
    assert r == 0.0 and np.array_equal(m, case["query_pose"]) and c[0, 0] == 500.0 and c[1, 1] == 500.0
    assert c[2, 2] == 4 * mapper["coarse_angle_resolution"] ** 2


def test_order_dependent_raster_matches_and_depends_on_order():
    case = synth.make_sequential_case(3, buffer_len=4)
    pm, gm = H.port_matcher(H.MAPPER_SEQ, H.GRID_SEQ_YAML), H.gpu_matcher(H.MAPPER_SEQ, H.GRID_SEQ_YAML)
    pq, gq = H.port_scans(case["query_ranges"], case["query_pose"])[0], H.gpu_block(case["query_ranges"], case["query_pose"])
    grids = []
    for order in (slice(None), slice(None, None, -1)):
        pb = H.port_scans(case["base_ranges"][order], case["base_poses"][order])
        gb = H.gpu_block(case["base_ranges"][order], case["base_poses"][order])
        pm.raster(pq, pb); gm.raster(gq, gb)
        assert np.array_equal(pm.grid()["data"], gm.GetCorrelationGrid()["data"])
        grids.append(pm.grid()["data"])
    assert not np.array_equal(grids[0], grids[1])


@pytest.mark.parametrize("chain_len,nch", [(1, 40), (10, 12)])
@pytest.mark.parametrize("refine", [False, True])
def test_batch_sweep_vs_oracle(chain_len, nch, refine):
    sw = synth.make_loop_sweep(3, n_queries=2, n_chains=nch, chain_len=chain_len, inf_frac=0.02)
    pm, gm = H.port_matcher(H.MAPPER_LOOP, H.GRID_LOOP), H.gpu_matcher(H.MAPPER_LOOP, H.GRID_LOOP)
    pc, pq = H.port_scans(sw.cand_ranges, sw.cand_poses), H.port_scans(sw.query_ranges, sw.query_poses)
    gc, gq = H.gpu_block(sw.cand_ranges, sw.cand_poses), H.gpu_block(sw.query_ranges, sw.query_poses)
    exp = [pm.match(pq[q], pc[sw.chain_start[c]:sw.chain_start[c + 1]], False, refine) for q in range(2) for c in range(nch)]
    resp, mean, cov = gm.MatchScanBatch(gq, gc, sw.chain_start, None, False, refine)
    assert np.array_equal(resp, np.array([e[0] for e in exp]))
    assert np.array_equal(mean, np.array([e[1] for e in exp]))
    assert np.array_equal(cov, np.array([e[2] for e in exp]))
    # an explicit pair list gives the same rows, and matches one-by-one MatchScan calls
    pairs = (np.array([1, 0, 1]), np.array([nch - 1, 3, 0]))
    r2, m2, c2 = gm.MatchScanBatch(gq, gc, sw.chain_start, pairs, False, refine)
    for k, (q, c) in enumerate(zip(*pairs)):
        assert r2[k] == resp[q * nch + c] and np.array_equal(m2[k], mean[q * nch + c]) and np.array_equal(c2[k], cov[q * nch + c])


def test_batch_with_penalty_and_order_dependent_kernel():
    """The batched path with the sequential-matcher parameter set (penalised, smear 0.1 @ 0.01 m: the raster is
    order dependent) -- exercises the on-device greedy raster rule and the FP64 penalty."""
    sw = synth.make_loop_sweep(9, n_queries=1, n_chains=3, chain_len=3, radius=0.6, drift_xy=0.05, drift_th=0.01)
    pm, gm = H.port_matcher(H.MAPPER_SEQ, H.GRID_SEQ_YAML), H.gpu_matcher(H.MAPPER_SEQ, H.GRID_SEQ_YAML)
    pc, pq = H.port_scans(sw.cand_ranges, sw.cand_poses), H.port_scans(sw.query_ranges, sw.query_poses)
    gc, gq = H.gpu_block(sw.cand_ranges, sw.cand_poses), H.gpu_block(sw.query_ranges, sw.query_poses)
    exp = [pm.match(pq[0], pc[sw.chain_start[c]:sw.chain_start[c + 1]], True, True) for c in range(3)]
    resp, mean, cov = gm.MatchScanBatch(gq, gc, sw.chain_start, None, True, True)
    assert np.array_equal(resp, np.array([e[0] for e in exp]))
    assert np.array_equal(mean, np.array([e[1] for e in exp])) and np.array_equal(cov, np.array([e[2] for e in exp]))


def test_full_size_sweep_properties():
    """BASELINE cfg2 size (1 x 1000 candidates): properties that need no CPU oracle run --
    idempotence, agreement of the split (upload/run/fetch) and fused entry points, best key = max of the
    per-pair integer sums, and a sampled exact comparison with the oracle."""
    sw = synth.make_loop_sweep(21, n_queries=1, n_chains=1000, chain_len=1)
    gm = H.gpu_matcher(H.MAPPER_LOOP, H.GRID_LOOP)
    gc, gq = H.gpu_block(sw.cand_ranges, sw.cand_poses), H.gpu_block(sw.query_ranges, sw.query_poses)
    r1, m1, c1 = gm.MatchScanBatch(gq, gc, sw.chain_start, None, False, False)
    assert gm.batch_info()["fast"]
    gm.batch_upload(gq, gc, sw.chain_start, None, False)
    gm.batch_run(); gm.batch_run()
    r2, m2, c2 = gm.batch_fetch()
    assert np.array_equal(r1, r2) and np.array_equal(m1, m2) and np.array_equal(c1, c2)
    sums, idx, ties = gm.batch_best()
    assert np.array_equal(np.minimum(1.0, sums / (1081 * 100.0)), r1)
    assert (idx >= 0).all() and (idx < 41 * 41 * 21).all() and (ties >= 1).all()
    import torch
    keys = torch.zeros(1, dtype=torch.int64, device="cuda")
    gm.batch_reduce_keys(keys.data_ptr(), 5000)
    torch.cuda.synchronize()
    from slam_toolbox_b200 import sweep
    s, g = sweep.unpack_keys(keys.cpu().numpy())
    assert s[0] == sums.max() and g[0] == 5000 + int(np.argmax(sums))
    assert (r1 >= 0).all() and (r1 <= 1).all() and r1.max() > 0.5
    pm = H.port_matcher(H.MAPPER_LOOP, H.GRID_LOOP)
    pq = H.port_scans(sw.query_ranges, sw.query_poses)[0]
    for j in (0, 17, 500, 999, int(np.argmax(r1))):
        e = pm.match(pq, H.port_scans(sw.cand_ranges[j], sw.cand_poses[j]), False, False)
        assert e[0] == r1[j] and np.array_equal(e[1], m1[j]) and np.array_equal(e[2], c1[j])


@pytest.mark.parametrize("grid", [H.GRID_LOOP, (4.0, 0.05, 0.03, 6.0), (2.0, 0.05, 0.05, 8.0)])
def test_fast_and_generic_sweep_kernels_agree_with_the_oracle(grid):
    """The shared-memory fast path (parity sub-grids, word loads) against the generic kernel and the oracle.
    A short range threshold puts many beam windows partly / wholly outside the grid, which exercises the
    SLOW-beam rule (linear-index wrap, Mapper.cpp:1192-1197) next to the FAST lists."""
    sw = synth.make_loop_sweep(31, n_queries=2, n_chains=10, chain_len=2, inf_frac=0.02)
    pm, gm = H.port_matcher(H.MAPPER_LOOP, grid), H.gpu_matcher(H.MAPPER_LOOP, grid)
    pc, pq = H.port_scans(sw.cand_ranges, sw.cand_poses), H.port_scans(sw.query_ranges, sw.query_poses)
    gc, gq = H.gpu_block(sw.cand_ranges, sw.cand_poses), H.gpu_block(sw.query_ranges, sw.query_poses)
    exp = [pm.match(pq[q], pc[sw.chain_start[c]:sw.chain_start[c + 1]], False, False) for q in range(2) for c in range(10)]
    gm.set_option("sweep_kernel", 1)           # the single-CTA shared-memory kernel (the default is the tiled cluster kernel)
    legacy = gm.MatchScanBatch(gq, gc, sw.chain_start, None, False, False)
    assert gm.batch_info()["kernel"] == "fast", gm.batch_info()
    legacy_best = gm.batch_best()
    gm.set_option("sweep_kernel", 0)
    fast = gm.MatchScanBatch(gq, gc, sw.chain_start, None, False, False)
    info = gm.batch_info()
    assert info["kernel"] == "tile" and info["fast_descriptors"] > 0 and info["edge_beams"] >= 0, info      # a shared-memory path really ran
    for a, b in zip(fast, legacy):
        assert np.array_equal(a, b)
    for a, b in zip(gm.batch_best(), legacy_best):
        assert np.array_equal(a, b)
    if grid[3] < 12.0:
        assert info["edge_beams"] > 0, info    # short range threshold: windows that leave the grid are exercised
    fast_best = gm.batch_best()
    gm.set_option("no_beam_dedup", 1)          # one descriptor per beam instead of (descriptor, multiplicity)
    plain = gm.MatchScanBatch(gq, gc, sw.chain_start, None, False, False)
    for a, b in zip(fast, plain):
        assert np.array_equal(a, b)
    gm.set_option("no_beam_dedup", 0)
    gm.set_option("force_generic_sweep", 1)
    gen = gm.MatchScanBatch(gq, gc, sw.chain_start, None, False, False)
    gen_best = gm.batch_best()
    for a, b in zip(fast, gen):
        assert np.array_equal(a, b)
    for a, b in zip(fast_best, gen_best):
        assert np.array_equal(a, b)
    assert np.array_equal(fast[0], np.array([e[0] for e in exp]))
    assert np.array_equal(fast[1], np.array([e[1] for e in exp])) and np.array_equal(fast[2], np.array([e[2] for e in exp]))


def test_sharded_sweep_keys_equal_the_unsharded_reduction():
    """cfg5 shape at small size: Q queries x C candidate chains, candidates sharded over "ranks" (here: two handles on one
    GPU, the reduction done with torch.maximum exactly as all_reduce(MAX) would). The per-query winner must not depend on
    the sharding: highest integer sum, ties to the lowest global candidate id."""
    import torch
    from slam_toolbox_b200 import sweep
    Q, Cn = 4, 48
    sw = synth.make_loop_sweep(41, n_queries=Q, n_chains=Cn, chain_len=1)
    gq = H.gpu_block(sw.query_ranges, sw.query_poses)

    def keys_for(lo, hi, offset):
        gm = H.gpu_matcher(H.MAPPER_LOOP, H.GRID_LOOP)
        gc = H.gpu_block(sw.cand_ranges[lo:hi], sw.cand_poses[lo:hi])
        gm.batch_upload(gq, gc, np.arange(hi - lo + 1, dtype=np.int32), None, False)
        gm.batch_run()
        k = torch.zeros(Q, dtype=torch.int64, device="cuda")
        gm.batch_reduce_keys(k.data_ptr(), offset)
        torch.cuda.synchronize()
        sums, _, _ = gm.batch_best()
        return k, sums.reshape(Q, hi - lo), gm.batch_fetch()

    full, sums, (resp, mean, cov) = keys_for(0, Cn, 0)
    parts, results, ranges = [], [], []
    for r in range(3):
        lo, hi = sweep.shard_range(Cn, 3, r)
        k, _, res = keys_for(lo, hi, lo)
        parts.append(k); results.append(res); ranges.append((lo, hi))
    merged = torch.maximum(torch.maximum(parts[0], parts[1]), parts[2])
    assert torch.equal(merged, full)
    s, g = sweep.unpack_keys(full.cpu().numpy())
    assert np.array_equal(s, sums.max(axis=1)) and np.array_equal(g, sums.argmax(axis=1))
    # winners exchange: each "rank" contributes the rows it owns, the sum is the unsharded winner's result
    table = np.zeros((Q, 13))
    for (lo, hi), (rr, mm, cc) in zip(ranges, results):
        pq = np.repeat(np.arange(Q), hi - lo)
        pc = np.tile(np.arange(lo, hi), Q)
        table += sweep.winners_payload(merged.cpu().numpy(), lo, hi, pq, pc, rr, mm, cc)
    win = np.arange(Q) * Cn + g
    assert np.array_equal(table[:, 0], resp[win]) and np.array_equal(table[:, 1:4], mean[win])
    assert np.array_equal(table[:, 4:], cov[win].reshape(Q, 9))


@pytest.mark.parametrize("n_beams,angle_min_deg,inc_deg", [(360, -180.0, 1.0), (1440, -180.0, 0.25), (721, -90.0, 0.25), (50, -25.0, 1.0)])
def test_other_lasers_single_and_batch(n_beams, angle_min_deg, inc_deg):
    """Lasers other than the 1081-beam one of the BASELINE configs (a 360-beam 1-degree lidar, a 1440-beam full circle,
    a 180-degree scanner, a 50-beam toy): single match (coarse + fine) and the batched sweep against the oracle."""
    from oracle import karto_port as P
    amin, ainc = np.radians(angle_min_deg), np.radians(inc_deg)
    laser = api.LaserRangeFinder(minimum_angle=amin, maximum_angle=amin + (n_beams - 1) * ainc, angular_resolution=ainc)
    rng = np.random.default_rng(n_beams)
    world = synth.make_world(77)
    qtrue = synth.free_pose(world, rng)
    cposes = synth.poses_near(world, qtrue[:2], 2.0, 9, rng)
    qpose = qtrue + np.array([0.2, -0.15, 0.05])
    qr = synth.noisy(synth.raycast(world, qtrue, n_beams, amin, ainc), rng, inf_frac=0.02)
    cr = synth.noisy(synth.raycast(world, cposes, n_beams, amin, ainc), rng, inf_frac=0.02)
    pq = P.PortScan(qr[0], qpose, amin, ainc)
    pc = [P.PortScan(r, p, amin, ainc) for r, p in zip(cr, cposes)]
    gq, gc = api.ScanBlock(qr, qpose[None, :], laser), api.ScanBlock(cr, cposes, laser)
    for mapper, grid, pen, refine in ((H.MAPPER_LOOP, H.GRID_LOOP, False, False), (H.MAPPER_SEQ, H.GRID_SEQ, True, True)):
        pm, gm = H.port_matcher(mapper, grid), H.gpu_matcher(mapper, grid)
        assert same(pm.match(pq, pc[:3], pen, refine), gm.MatchScan(gq, api.ScanBlock(cr[:3], cposes[:3], laser), pen, refine))
    pm, gm = H.port_matcher(H.MAPPER_LOOP, H.GRID_LOOP), H.gpu_matcher(H.MAPPER_LOOP, H.GRID_LOOP)
    chain_start = np.array([0, 1, 2, 3, 6, 9], dtype=np.int32)      # three single scans and two chains of three
    exp = [pm.match(pq, pc[chain_start[c]:chain_start[c + 1]], False, False) for c in range(5)]
    resp, mean, cov = gm.MatchScanBatch(gq, gc, chain_start, None, False, False)
    for c in range(5):
        assert same(exp[c], (resp[c], mean[c], cov[c])), c
    assert gm.batch_info()["fast"]
