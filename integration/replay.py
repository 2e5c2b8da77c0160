"""cfg3 replay: the reference's own karto::Mapper::Process over a sequence of posed scans, through one of
the two integration libraries built by integration/Makefile:

  libreplay_ref.so   reference CPU ScanMatcher            + GPU ScanSolver adapter (B200Solver)
  libreplay_b200.so  GPU ScanMatcher (link-time seam)     + GPU ScanSolver adapter

Each run happens in its own process (both libraries define the same karto symbols):
    python integration/replay.py <ref|b200> <in.npz> <out.npz>
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
_DP = C.POINTER(C.c_double)

# slam_toolbox's shipped parameter set (config/mapper_params_online_sync.yaml:39-74), with Karto's default
# smear so that the sequential raster is order independent; BASELINE cfg3
YAML_PARAMS = dict(
    use_scan_matching=1, use_scan_barycenter=1, minimum_travel_distance=0.5, minimum_travel_heading=0.5, scan_buffer_size=10,
    scan_buffer_maximum_scan_distance=10.0, link_match_minimum_response_fine=0.1, link_scan_maximum_distance=1.5,
    loop_search_maximum_distance=3.0, do_loop_closing=1, loop_match_minimum_chain_size=10, loop_match_maximum_variance_coarse=3.0,
    loop_match_minimum_response_coarse=0.35, loop_match_minimum_response_fine=0.45,
    correlation_search_space_dimension=0.5, correlation_search_space_resolution=0.01, correlation_search_space_smear_deviation=0.1,
    loop_search_space_dimension=8.0, loop_search_space_resolution=0.05, loop_search_space_smear_deviation=0.03,
    distance_variance_penalty=0.5, angle_variance_penalty=1.0, fine_search_angle_offset=0.00349, coarse_search_angle_offset=0.349,
    coarse_angle_resolution=0.0349, minimum_angle_penalty=0.9, minimum_distance_penalty=0.5, use_response_expansion=1)


def library(which: str) -> str:
    return os.path.join(HERE, "_build", f"libreplay_{which}.so")


def available() -> bool:
    return os.path.exists(library("ref")) and os.path.exists(library("b200"))


def run_inprocess(which: str, ranges: np.ndarray, odom: np.ndarray, params: dict, laser: dict, use_solver: bool = True,
                  map_resolution: float = 0.0):
    L = C.CDLL(library(which))
    L.krep_create.restype = C.c_void_p
    L.krep_create.argtypes = [C.c_int]
    L.krep_set.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
    L.krep_process.argtypes = [C.c_void_p, _DP, C.c_int, _DP, C.c_int]
    L.krep_num_scans.argtypes = [C.c_void_p]
    L.krep_poses.argtypes = [C.c_void_p, _DP]
    L.krep_stats.argtypes = [C.c_void_p, _DP]
    L.krep_init_laser.argtypes = [C.c_double] * 6
    L.krep_init_laser(laser["min_angle"], laser["max_angle"], laser["ang_res"], laser["min_range"], laser["max_range"],
                      laser["range_threshold"])
    h = L.krep_create(int(use_solver))
    for k, v in params.items():
        if L.krep_set(h, k.encode(), float(v)) != 0:
            raise KeyError(k)
    ranges = np.ascontiguousarray(ranges, dtype=np.float64)
    odom = np.ascontiguousarray(odom, dtype=np.float64)
    kept = []
    for i in range(len(ranges)):
        if L.krep_process(h, ranges[i].ctypes.data_as(_DP), ranges.shape[1], odom[i].ctypes.data_as(_DP), i):
            kept.append(i)
    n = L.krep_num_scans(h)
    poses = np.zeros((n, 3))
    L.krep_poses(h, poses.ctypes.data_as(_DP))
    st = np.zeros(5)
    L.krep_stats(h, st.ctypes.data_as(_DP))
    matches = 0
    if which == "b200":
        L.b200_shim_match_calls.restype = C.c_long
        matches = int(L.b200_shim_match_calls())
    L.krep_solver_graph.argtypes = [C.c_void_p, C.POINTER(C.c_int), _DP, C.c_int]
    gids = np.zeros(max(n, 1), dtype=np.int32)
    gposes = np.zeros((max(n, 1), 3))
    gn = L.krep_solver_graph(h, gids.ctypes.data_as(C.POINTER(C.c_int)), gposes.ctypes.data_as(_DP), n) if use_solver else 0
    order = np.argsort(gids[:min(gn, n)])
    out = dict(graph_nodes=gn, graph_ids=gids[:min(gn, n)][order], graph_poses=gposes[:min(gn, n)][order], poses=poses, kept=np.array(kept),
               process_seconds=st[0], solver_computes=int(st[1]), solver_ms=st[2], edges=int(st[3]), scans=int(st[4]), match_calls=matches)
    if map_resolution:
        # the map-publish step over all processed scans: reference CPU build and the b200og binding, same process
        L.krep_occupancy.restype = C.c_double
        L.krep_occupancy.argtypes = [C.c_void_p, C.c_double, C.c_int, C.POINTER(C.c_int), _DP, C.POINTER(C.c_uint8), C.c_long]
        for tag, gpu in (("cpu", 0), ("gpu", 1)):
            info = (C.c_int * 3)()
            off = np.zeros(2)
            sec = L.krep_occupancy(h, map_resolution, gpu, info, off.ctypes.data_as(_DP), None, 0)   # sizes (and warm-up)
            cells = np.zeros((max(info[1], 0), max(info[2], 0)), dtype=np.uint8)
            if sec >= 0:
                sec = L.krep_occupancy(h, map_resolution, gpu, info, off.ctypes.data_as(_DP), cells.ctypes.data_as(C.POINTER(C.c_uint8)),
                                       cells.size)
            out[f"map_{tag}_seconds"] = sec
            out[f"map_{tag}_dims"] = np.array([info[0], info[1], info[2]])
            out[f"map_{tag}_offset"] = off
            out[f"map_{tag}_cells"] = cells
    return out


def run(which: str, ranges, odom, params=None, laser=None, use_solver=True, tmpdir=None, map_resolution=0.0):
    """Runs the replay in a fresh process and returns its result dict."""
    import tempfile
    params = params or YAML_PARAMS
    laser = laser or dict(min_angle=math.radians(-135), max_angle=math.radians(135), ang_res=math.radians(0.25), min_range=0.1,
                          max_range=30.0, range_threshold=12.0)
    d = tmpdir or tempfile.mkdtemp(prefix="replay_")
    fin, fout = os.path.join(d, f"in_{which}.npz"), os.path.join(d, f"out_{which}.npz")
    np.savez(fin, ranges=ranges, odom=odom, pkeys=np.array(list(params.keys())), pvals=np.array(list(params.values()), dtype=np.float64),
             lkeys=np.array(list(laser.keys())), lvals=np.array(list(laser.values()), dtype=np.float64), use_solver=int(use_solver),
             map_resolution=float(map_resolution))
    env = dict(os.environ)
    r = subprocess.run([sys.executable, os.path.abspath(__file__), which, fin, fout], capture_output=True, text=True, env=env)
    if r.returncode != 0:
        raise RuntimeError(f"replay {which} failed:\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}")
    z = np.load(fout)
    return {k: (z[k] if z[k].shape else z[k].item()) for k in z.files}


def make_trajectory(seed: int, n_scans: int, step: float = 0.5):
    """A wandering path with revisits through the synthetic world: posed scans + drifting odometry."""
    sys.path.insert(0, ROOT)
    from slam_toolbox_b200 import synth
    rng = np.random.default_rng(seed)
    world = synth.make_world(seed)
    # a loop through a few room centres, repeated, so that later passes revisit earlier ones
    start = synth.free_pose(world, rng)
    half = synth.chain_poses(world, start, max(8, n_scans // 2), rng, step=step)
    back = half[::-1].copy()
    back[:, 2] = synth.wrap(back[:, 2] + math.pi)
    traj = np.concatenate([half, back])[:n_scans]
    ranges = synth.noisy(synth.raycast(world, traj), rng)
    # odometry = truth + slowly accumulating drift
    drift = np.cumsum(np.column_stack([rng.normal(0, 0.004, (len(traj), 2)), rng.normal(0, 0.0015, len(traj))]), axis=0)
    odom = traj + drift
    return ranges, odom, traj


def lifecycle_inprocess():
    """Adapter hardening checks that need the integration library in this process: option mapping of B200Solver (the ceres_* keys
    CeresSolver::Configure reads) and release of the matchers' device state on Mapper::Reset / destruction."""
    L = C.CDLL(library("b200"))
    L.krep_create.restype = C.c_void_p
    L.krep_create.argtypes = [C.c_int]
    L.krep_process.argtypes = [C.c_void_p, _DP, C.c_int, _DP, C.c_int]
    L.krep_set.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
    L.krep_solver_configure.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
    L.krep_reset_mapper.argtypes = [C.c_void_p]
    L.krep_destroy.argtypes = [C.c_void_p]
    L.b200_shim_live_handles.restype = C.c_long
    L.krep_init_laser.argtypes = [C.c_double] * 6
    L.krep_init_laser(math.radians(-135), math.radians(135), math.radians(0.25), 0.1, 30.0, 12.0)
    ranges, odom, _ = make_trajectory(3, 24)
    h = L.krep_create(1)
    for k, v in YAML_PARAMS.items():
        L.krep_set(h, k.encode(), float(v))
    out = {"configure": {}}
    for key, val in (("ceres_loss_function", "HuberLoss"), ("ceres_loss_function", "CauchyLoss"), ("ceres_loss_function", "None"),
                     ("ceres_loss_function", "Bogus"), ("ceres_trust_strategy", "LEVENBERG_MARQUARDT"), ("ceres_trust_strategy", "DOGLEG"),
                     ("ceres_linear_solver", "SPARSE_NORMAL_CHOLESKY"), ("ceres_preconditioner", "SCHUR_JACOBI"), ("no_such_key", "1")):
        out["configure"][f"{key}={val}"] = int(L.krep_solver_configure(h, key.encode(), val.encode()))
    live = [int(L.b200_shim_live_handles())]

    def feed(lo, hi):
        for i in range(lo, hi):
            L.krep_process(h, np.ascontiguousarray(ranges[i]).ctypes.data_as(_DP), ranges.shape[1], np.ascontiguousarray(odom[i]).ctypes.data_as(_DP), i)
    feed(0, 12)
    live.append(int(L.b200_shim_live_handles()))     # sequential + loop matcher
    L.krep_reset_mapper(h)
    live.append(int(L.b200_shim_live_handles()))     # Mapper::Reset deleted both
    feed(12, 24)
    live.append(int(L.b200_shim_live_handles()))
    L.krep_destroy(h)
    live.append(int(L.b200_shim_live_handles()))
    out["live_handles"] = live
    return out


if __name__ == "__main__":
    if sys.argv[1] == "lifecycle":
        import json
        devnull = os.open(os.devnull, os.O_WRONLY)
        saved = os.dup(1)
        os.dup2(devnull, 1)
        try:
            res = lifecycle_inprocess()
        finally:
            os.dup2(saved, 1)
        print(json.dumps(res))
        sys.exit(0)
    which, fin, fout = sys.argv[1], sys.argv[2], sys.argv[3]
    z = np.load(fin)
    params = {str(k): float(v) for k, v in zip(z["pkeys"], z["pvals"])}
    laser = {str(k): float(v) for k, v in zip(z["lkeys"], z["lvals"])}
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(1)
    os.dup2(devnull, 1)   # the reference prints progress to stdout
    try:
        out = run_inprocess(which, z["ranges"], z["odom"], params, laser, bool(int(z["use_solver"])), float(z["map_resolution"]))
    finally:
        os.dup2(saved, 1)
    np.savez(fout, **out)
