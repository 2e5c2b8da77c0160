"""Pins the plain-C oracle (oracle/karto_port.c) against the reference itself: the unmodified
karto_sdk compiled into oracle/_ref/libkarto_ref.so (present in the build container; travels to the
GPU box as a prebuilt .so) and against the committed golden fixtures generated from it."""
import hashlib
import math
import os

import numpy as np
import pytest

import helpers as H
from oracle import karto_port as P
from oracle import karto_ref as R
from slam_toolbox_b200 import synth

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "matcher_golden.npz")
needs_ref = pytest.mark.skipif(not R.available(), reason="oracle/_ref/libkarto_ref.so not built")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def golden_cases():
    z = np.load(GOLDEN)
    names = sorted({k.split("/")[0] for k in z.files})
    return z, names


CONFIG = {"seq_k03": (H.MAPPER_SEQ, H.GRID_SEQ), "seq_yaml_inf": (H.MAPPER_SEQ, H.GRID_SEQ_YAML),
          "loop_chain5": (H.MAPPER_LOOP, H.GRID_LOOP), "loop_refine": (H.MAPPER_LOOP, H.GRID_LOOP),
          "small": (H.MAPPER_LOOP, H.GRID_SMALL)}


@pytest.mark.parametrize("name", sorted(CONFIG))
def test_port_matches_golden(name):
    z, _ = golden_cases()
    mapper, grid = CONFIG[name]
    pm = H.port_matcher(mapper, grid)
    base = H.port_scans(z[f"{name}/base_ranges"], z[f"{name}/base_poses"])
    q = H.port_scans(z[f"{name}/query_ranges"], z[f"{name}/query_pose"])[0]
    pen, refine = (bool(v) for v in z[f"{name}/flags"])
    resp, mean, cov = pm.match(q, base, pen, refine)
    assert resp == z[f"{name}/response"][0]
    assert np.array_equal(mean, z[f"{name}/mean"])
    assert np.array_equal(cov, z[f"{name}/cov"])
    assert np.array_equal(pm.kernel(), z[f"{name}/kernel"])
    pm.raster(q, base)
    assert sha(pm.grid()["data"]) == z[f"{name}/grid_sha"][0]
    off = pm.offsets(q, z[f"{name}/query_pose"][2], mapper["coarse_search_angle_offset"], mapper["coarse_angle_resolution"])
    assert sha(off) == z[f"{name}/offsets_sha"][0]
    so, sr = H.coarse_search(grid)
    _, _, _, vol = pm.correlate(q, z[f"{name}/query_pose"], so, sr, mapper["coarse_search_angle_offset"],
                                mapper["coarse_angle_resolution"], False, False)
    assert sha(vol) == z[f"{name}/volume_sha"][0]
    assert int(vol.argmax()) == z[f"{name}/volume_argmax"][0] and int(vol.max()) == z[f"{name}/volume_argmax"][1]


@needs_ref
@pytest.mark.parametrize("seed", [0, 1, 2])
@pytest.mark.parametrize("cfg", ["seq", "seq_yaml", "loop"])
def test_port_vs_reference_match(seed, cfg):
    mapper, grid = {"seq": (H.MAPPER_SEQ, H.GRID_SEQ), "seq_yaml": (H.MAPPER_SEQ, H.GRID_SEQ_YAML),
                    "loop": (H.MAPPER_LOOP, H.GRID_LOOP)}[cfg]
    case = synth.make_sequential_case(100 + seed, buffer_len=4, inf_frac=0.03 * (seed % 2), nan_frac=0.01 * (seed == 2))
    rm, pm = H.ref_matcher(mapper, grid), H.port_matcher(mapper, grid)
    rb, pb = H.ref_scans(case["base_ranges"], case["base_poses"]), H.port_scans(case["base_ranges"], case["base_poses"])
    rq, pq = H.ref_scans(case["query_ranges"], case["query_pose"], 99)[0], H.port_scans(case["query_ranges"], case["query_pose"])[0]
    assert np.array_equal(rq.points(), pq.points, equal_nan=True)
    for pen, refine in ((True, True), (False, False), (False, True)):
        a, b = rm.match(rq, rb, pen, refine), pm.match(pq, pb, pen, refine)
        assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    g1, g2 = rm.grid(), pm.grid()
    assert np.array_equal(g1["data"], g2["data"]) and g1["offset"] == g2["offset"]
    assert (g1["width"], g1["stride"], g1["roi"], g1["kernel_size"]) == (g2["width"], g2["stride"], g2["roi"], g2["kernel_size"])
    assert np.array_equal(rm.kernel(), pm.kernel())
    o1 = rm.offsets(rq, case["query_pose"][2] + 0.01, mapper["coarse_search_angle_offset"], mapper["coarse_angle_resolution"])
    o2 = pm.offsets(pq, case["query_pose"][2] + 0.01, mapper["coarse_search_angle_offset"], mapper["coarse_angle_resolution"])
    assert np.array_equal(o1, o2)
    vp = case["query_pose"][:2] + 0.3
    for r_s, p_s in zip(rb, pb):
        assert np.array_equal(rm.find_valid_points(r_s, vp), P.find_valid_points(p_s, vp), equal_nan=True)


@needs_ref
def test_port_vs_reference_edge_cases():
    mapper, grid = H.MAPPER_LOOP, H.GRID_SMALL
    rm, pm = H.ref_matcher(mapper, grid), H.port_matcher(mapper, grid)
    case = synth.make_sequential_case(7, buffer_len=2)
    rq, pq = H.ref_scans(case["query_ranges"], case["query_pose"], 5)[0], H.port_scans(case["query_ranges"], case["query_pose"])[0]
    # no base scans at all: zero response everywhere, every pose ties, response expansion kicks in
    a, b = rm.match(rq, [], True, True), pm.match(pq, [], True, True)
    assert a[0] == b[0] == 0.0 and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    # all-invalid query readings
    bad = np.full_like(case["query_ranges"], np.inf)
    rq2, pq2 = H.ref_scans(bad, case["query_pose"], 6)[0], H.port_scans(bad, case["query_pose"])[0]
    rb, pb = H.ref_scans(case["base_ranges"], case["base_poses"]), H.port_scans(case["base_ranges"], case["base_poses"])
    a, b = rm.match(rq2, rb, False, False), pm.match(pq2, pb, False, False)
    assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    # query far away from the base scans (nothing overlaps)
    far = case["query_pose"] + np.array([500.0, -300.0, 1.0])
    rq3, pq3 = H.ref_scans(case["query_ranges"], far, 7)[0], H.port_scans(case["query_ranges"], far)[0]
    a, b = rm.match(rq3, rb, True, False), pm.match(pq3, pb, True, False)
    assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])


@needs_ref
def test_raster_order_dependence_is_reproduced():
    """SURVEY.md 7 hard part 2: with smear 0.1 @ 0.01 m the grid depends on base-scan order."""
    case = synth.make_sequential_case(3, buffer_len=4)
    rm, pm = H.ref_matcher(H.MAPPER_SEQ, H.GRID_SEQ_YAML), H.port_matcher(H.MAPPER_SEQ, H.GRID_SEQ_YAML)
    rq, pq = H.ref_scans(case["query_ranges"], case["query_pose"], 9)[0], H.port_scans(case["query_ranges"], case["query_pose"])[0]
    rb, pb = H.ref_scans(case["base_ranges"], case["base_poses"]), H.port_scans(case["base_ranges"], case["base_poses"])
    grids = []
    for order in (slice(None), slice(None, None, -1)):
        rm.raster(rq, rb[order]); pm.raster(pq, pb[order])
        assert np.array_equal(rm.grid()["data"], pm.grid()["data"])
        grids.append(pm.grid()["data"])
    assert not np.array_equal(grids[0], grids[1])


def test_create_rejects_what_the_reference_rejects():
    for bad in (dict(search_size=-1.0), dict(resolution=0.0), dict(smear_deviation=-0.1), dict(range_threshold=0.0),
                dict(smear_deviation=1.0), dict(smear_deviation=0.001)):
        kw = dict(search_size=1.0, resolution=0.05, smear_deviation=0.03, range_threshold=6.0, coarse_search_angle_offset=0.3,
                  coarse_angle_resolution=0.03, fine_search_angle_offset=0.003, distance_variance_penalty=0.25,
                  angle_variance_penalty=1.0, minimum_distance_penalty=0.5, minimum_angle_penalty=0.9, use_response_expansion=0)
        kw.update(bad)
        with pytest.raises(ValueError):
            P.PortMatcher(**kw)
