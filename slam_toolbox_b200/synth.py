"""Synthetic inputs for the scan-matching / pose-graph hot path.

Everything here is *workload generation* for tests and bench.py: a Manhattan world of
axis-aligned wall segments on a 0.5 m lattice, exact ray casting for a 1081-beam laser
(-135 deg .. +135 deg @ 0.25 deg, like the reference survey's LaserRangeFinder_Custom set-up),
posed scans, loop-closure candidate sets and a Manhattan-world SE(2) pose graph.  Nothing in
this file is on the product path.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

N_BEAMS = 1081
ANGLE_MIN = math.radians(-135.0)
ANGLE_MAX = math.radians(135.0)
ANGLE_INC = math.radians(0.25)
RANGE_MIN = 0.1
RANGE_MAX = 30.0


@dataclass
class World:
    vert: np.ndarray   # (V,3) x, y0, y1   wall segments x = const
    horz: np.ndarray   # (H,3) y, x0, x1   wall segments y = const
    rooms: np.ndarray  # (R,4) x0, y0, x1, y1
    size: float

    @property
    def n_segments(self) -> int:
        return len(self.vert) + len(self.horz)


def make_world(seed: int, size: float = 36.0, min_room: float = 4.0, max_room: float = 12.0) -> World:
    """BSP split of [0,size]^2 into rooms with sides in [min_room, max_room] on a 0.5 m lattice;
    each interior wall gets one 1 m door."""
    rng = np.random.default_rng(seed)
    vert, horz, rooms = [], [], []

    def lattice(a: float, b: float) -> float:
        k0, k1 = int(math.ceil(a * 2 - 1e-9)), int(math.floor(b * 2 + 1e-9))
        return 0.5 * int(rng.integers(k0, k1 + 1))

    def split(x0, y0, x1, y1):
        w, h = x1 - x0, y1 - y0
        can_x, can_y = w >= 2 * min_room, h >= 2 * min_room
        must = w > max_room or h > max_room
        if not (can_x or can_y) or (not must and rng.random() < 0.5):
            rooms.append((x0, y0, x1, y1))
            return
        if can_x and (not can_y or w >= h):
            c = lattice(x0 + min_room, x1 - min_room)
            d = lattice(y0 + 0.5, y1 - 1.5)
            if d - y0 > 1e-9:
                vert.append((c, y0, d))
            if y1 - (d + 1.0) > 1e-9:
                vert.append((c, d + 1.0, y1))
            split(x0, y0, c, y1)
            split(c, y0, x1, y1)
        else:
            c = lattice(y0 + min_room, y1 - min_room)
            d = lattice(x0 + 0.5, x1 - 1.5)
            if d - x0 > 1e-9:
                horz.append((c, x0, d))
            if x1 - (d + 1.0) > 1e-9:
                horz.append((c, d + 1.0, x1))
            split(x0, y0, x1, c)
            split(x0, c, x1, y1)

    split(0.0, 0.0, size, size)
    vert += [(0.0, 0.0, size), (size, 0.0, size)]
    horz += [(0.0, 0.0, size), (size, 0.0, size)]
    return World(np.array(vert, dtype=np.float64), np.array(horz, dtype=np.float64),
                 np.array(rooms, dtype=np.float64), size)


def raycast(world: World, poses: np.ndarray, n_beams: int = N_BEAMS, angle_min: float = ANGLE_MIN,
            angle_inc: float = ANGLE_INC, range_max: float = RANGE_MAX, chunk: int = 16) -> np.ndarray:
    """Exact ranges (P, n_beams) for sensor poses (P,3) against the axis-aligned segments."""
    poses = np.atleast_2d(np.asarray(poses, dtype=np.float64))
    out = np.empty((len(poses), n_beams), dtype=np.float64)
    beam = angle_min + np.arange(n_beams) * angle_inc
    vx, vy0, vy1 = world.vert[:, 0], world.vert[:, 1], world.vert[:, 2]
    hy, hx0, hx1 = world.horz[:, 0], world.horz[:, 1], world.horz[:, 2]
    for s in range(0, len(poses), chunk):
        p = poses[s:s + chunk]
        ang = p[:, 2:3] + beam[None, :]
        dx, dy = np.cos(ang)[..., None], np.sin(ang)[..., None]       # (c,n,1)
        ox, oy = p[:, 0][:, None, None], p[:, 1][:, None, None]
        with np.errstate(divide="ignore", invalid="ignore"):
            tv = (vx[None, None, :] - ox) / dx
            yv = oy + tv * dy
            okv = (tv > 1e-9) & (yv >= vy0 - 1e-12) & (yv <= vy1 + 1e-12)
            tv = np.where(okv, tv, np.inf)
            th = (hy[None, None, :] - oy) / dy
            xh = ox + th * dx
            okh = (th > 1e-9) & (xh >= hx0 - 1e-12) & (xh <= hx1 + 1e-12)
            th = np.where(okh, th, np.inf)
        r = np.minimum(tv.min(axis=2), th.min(axis=2))
        out[s:s + chunk] = np.minimum(r, range_max)
    return out


def free_pose(world: World, rng: np.random.Generator, margin: float = 0.6) -> np.ndarray:
    r = world.rooms[int(rng.integers(len(world.rooms)))]
    return np.array([rng.uniform(r[0] + margin, r[2] - margin), rng.uniform(r[1] + margin, r[3] - margin),
                     rng.uniform(-math.pi, math.pi)])


def wall_clearance(world: World, xy: np.ndarray) -> np.ndarray:
    """Distance from points (N,2) to the closest wall segment."""
    x, y = xy[:, 0:1], xy[:, 1:2]
    dv = np.hypot(x - world.vert[None, :, 0], np.clip(y, world.vert[None, :, 1], world.vert[None, :, 2]) - y)
    dh = np.hypot(np.clip(x, world.horz[None, :, 1], world.horz[None, :, 2]) - x, y - world.horz[None, :, 0])
    return np.minimum(dv.min(axis=1), dh.min(axis=1))


def poses_near(world: World, center_xy, radius: float, n: int, rng: np.random.Generator,
               margin: float = 0.4) -> np.ndarray:
    """n sensor poses uniformly within `radius` of center_xy, at least `margin` from any wall."""
    out = np.empty((0, 3))
    c = np.asarray(center_xy, dtype=np.float64)
    while len(out) < n:
        m = 2 * (n - len(out)) + 16
        rr = radius * np.sqrt(rng.random(m))
        aa = rng.uniform(-math.pi, math.pi, m)
        xy = c[None, :] + np.stack([rr * np.cos(aa), rr * np.sin(aa)], axis=1)
        ok = (xy[:, 0] > margin) & (xy[:, 0] < world.size - margin) & (xy[:, 1] > margin) & (xy[:, 1] < world.size - margin)
        ok &= wall_clearance(world, xy) >= margin
        th = rng.uniform(-math.pi, math.pi, m)
        out = np.concatenate([out, np.column_stack([xy, th])[ok]])
    return out[:n]


def noisy(ranges: np.ndarray, rng: np.random.Generator, sigma: float = 0.01, inf_frac: float = 0.0,
          nan_frac: float = 0.0) -> np.ndarray:
    r = ranges + rng.normal(0.0, sigma, ranges.shape)
    r = np.clip(r, 0.0, RANGE_MAX)
    if inf_frac > 0:
        r = np.where(rng.random(r.shape) < inf_frac, np.inf, r)
    if nan_frac > 0:
        r = np.where(rng.random(r.shape) < nan_frac, np.nan, r)
    return r


def chain_poses(world: World, start: np.ndarray, n: int, rng: np.random.Generator, step: float = 0.5) -> np.ndarray:
    """A short trajectory of n poses, `step` apart, staying clear of walls (running-scan chain)."""
    poses = [np.array(start, dtype=np.float64)]
    heading = start[2]
    tries = 0
    while len(poses) < n:
        h = heading + rng.normal(0, 0.15)
        nxt = poses[-1][:2] + step * np.array([math.cos(h), math.sin(h)])
        ok = 0.4 < nxt[0] < world.size - 0.4 and 0.4 < nxt[1] < world.size - 0.4 and \
            wall_clearance(world, nxt[None, :])[0] > 0.4
        if ok:
            poses.append(np.array([nxt[0], nxt[1], h]))
            heading = h
            tries = 0
        else:
            heading = rng.uniform(-math.pi, math.pi)
            tries += 1
            if tries > 200:
                poses.append(poses[-1].copy())
    return np.array(poses)


@dataclass
class LoopSweep:
    """cfg2 / cfg5 shaped input: query scans and candidate chains."""
    query_ranges: np.ndarray          # (Q, n)
    query_poses: np.ndarray           # (Q, 3)  believed (drifted) sensor poses
    query_true: np.ndarray            # (Q, 3)
    cand_ranges: np.ndarray           # (S, n)  all candidate scans
    cand_poses: np.ndarray            # (S, 3)  corrected sensor poses
    chain_start: np.ndarray           # (C+1,)  chain j = scans [chain_start[j], chain_start[j+1])
    world: World = field(repr=False, default=None)


def make_loop_sweep(seed: int, n_queries: int = 1, n_chains: int = 1000, chain_len: int = 1,
                    drift_xy: float = 0.5, drift_th: float = 0.08, radius: float = 3.0,
                    inf_frac: float = 0.0, world: World | None = None) -> LoopSweep:
    rng = np.random.default_rng(seed)
    world = world or make_world(seed)
    qtrue = np.array([free_pose(world, rng) for _ in range(n_queries)])
    qr = noisy(raycast(world, qtrue), rng, inf_frac=inf_frac)
    qpose = qtrue + np.column_stack([rng.normal(0, drift_xy, (n_queries, 2)), rng.normal(0, drift_th, n_queries)])
    # candidates live around the first query's true position (one revisit area per sweep)
    starts = poses_near(world, qtrue[0, :2], radius, n_chains, rng)
    if chain_len == 1:
        cposes = starts
    else:
        cposes = np.concatenate([chain_poses(world, s, chain_len, rng) for s in starts])
    cr = noisy(raycast(world, cposes), rng, inf_frac=inf_frac)
    chain_start = np.arange(0, n_chains * chain_len + 1, chain_len, dtype=np.int32)
    return LoopSweep(qr, qpose, qtrue, cr, cposes, chain_start, world)


def make_sequential_case(seed: int, buffer_len: int = 10, odo_xy: float = 0.03, odo_th: float = 0.01,
                         inf_frac: float = 0.0, nan_frac: float = 0.0):
    """cfg1: one query against a running buffer of `buffer_len` scans spaced 0.5 m."""
    rng = np.random.default_rng(seed)
    world = make_world(seed)
    traj = chain_poses(world, free_pose(world, rng), buffer_len + 1, rng)
    ranges = noisy(raycast(world, traj), rng, inf_frac=inf_frac, nan_frac=nan_frac)
    qtrue = traj[-1]
    qpose = qtrue + np.array([rng.normal(0, odo_xy), rng.normal(0, odo_xy), rng.normal(0, odo_th)])
    return dict(world=world, base_ranges=ranges[:-1], base_poses=traj[:-1], query_ranges=ranges[-1],
                query_pose=qpose, query_true=qtrue)


def make_mapping_run(seed: int, n_scans: int = 24, step: float = 0.5, inf_frac: float = 0.02, nan_frac: float = 0.005,
                     odd_readings: bool = True, world: World | None = None):
    """Input of the map-publish step (OccupancyGrid::CreateFromScans over all processed scans): a trajectory of
    n_scans sensor poses `step` apart with their scans.  odd_readings sprinkles the values AddScan treats specially
    (Karto.h:6158-6171): below the minimum range, at / beyond the maximum range, exactly at the range threshold."""
    rng = np.random.default_rng(seed)
    world = world or make_world(seed)
    poses = chain_poses(world, free_pose(world, rng), n_scans, rng, step=step)
    ranges = noisy(raycast(world, poses), rng, inf_frac=inf_frac, nan_frac=nan_frac)
    if odd_readings and ranges.size:
        m = ranges.shape[1]
        for s in range(len(ranges)):
            k = rng.integers(0, m, 6)
            ranges[s, k[0]] = 0.05
            ranges[s, k[1]] = 0.1          # == minimum range: ignored (<=)
            ranges[s, k[2]] = 30.0         # == maximum range: ignored (>=)
            ranges[s, k[3]] = 12.0         # == range threshold: traced, no hit
            ranges[s, k[4]] = 12.0 - 5e-7  # inside the KT_TOLERANCE band: traced to the end point, no hit
            ranges[s, k[5]] = 25.0         # beyond the threshold: traced up to it
    return dict(world=world, ranges=ranges, poses=poses)


def wrap(a):
    """[-pi, pi) like solvers/ceres_utils.h:27-32 NormalizeAngle."""
    a = np.asarray(a, dtype=np.float64)
    return a - 2.0 * math.pi * np.floor((a + math.pi) / (2.0 * math.pi))


def make_pose_graph(seed: int, n_nodes: int = 10000, n_edges: int = 40000, lattice: int = 100,
                    sigma_xy: float = 0.05, sigma_th: float = 0.02, min_gap: int = 50):
    """cfg4: Manhattan-world SE(2) graph: random walk on a lattice (1 m steps, 90 deg
    turns), n_nodes-1 odometry edges + loop edges between nodes on the same/adjacent lattice site
    with index gap > min_gap.  Edge = (a, b, z = pose_b in frame a + noise, cov = diag(sigma^2) rotated
    like LinkInfo::Update, Mapper.h:174-188).  Initial guess = dead-reckoned odometry."""
    rng = np.random.default_rng(seed)
    dirs = np.array([[1, 0], [0, 1], [-1, 0], [0, -1]])
    cell = np.zeros((n_nodes, 2), dtype=np.int64)
    head = np.zeros(n_nodes, dtype=np.int64)
    cell[0] = lattice // 2
    for i in range(1, n_nodes):
        h = head[i - 1]
        r = rng.random()
        if r < 0.25:
            h = (h + 1) % 4
        elif r < 0.5:
            h = (h + 3) % 4
        nxt = cell[i - 1] + dirs[h]
        k = 0
        while not (0 <= nxt[0] < lattice and 0 <= nxt[1] < lattice):
            h = (h + 1) % 4
            nxt = cell[i - 1] + dirs[h]
            k += 1
        head[i] = h
        cell[i] = nxt
    truth = np.column_stack([cell[:, 0].astype(float), cell[:, 1].astype(float), wrap(head * (math.pi / 2))])

    def rel(pa, pb):
        c, s = np.cos(pa[:, 2]), np.sin(pa[:, 2])
        dx, dy = pb[:, 0] - pa[:, 0], pb[:, 1] - pa[:, 1]
        return np.column_stack([c * dx + s * dy, -s * dx + c * dy, wrap(pb[:, 2] - pa[:, 2])])

    ea = list(range(n_nodes - 1))
    eb = list(range(1, n_nodes))
    # loop edges: bucket nodes by lattice site
    site = {}
    for i in range(n_nodes):
        site.setdefault((int(cell[i, 0]), int(cell[i, 1])), []).append(i)
    pairs = set()
    want = n_edges - (n_nodes - 1)
    order = rng.permutation(n_nodes)
    nb = [(0, 0), (1, 0), (0, 1), (-1, 0), (0, -1)]
    rounds = 0
    while len(pairs) < want and rounds < 64:
        for i in order:
            if len(pairs) >= want:
                break
            dxy = nb[int(rng.integers(len(nb)))]
            lst = site.get((int(cell[i, 0]) + dxy[0], int(cell[i, 1]) + dxy[1]))
            if not lst:
                continue
            j = lst[int(rng.integers(len(lst)))]
            a, b = (int(i), int(j)) if i < j else (int(j), int(i))
            if b - a > min_gap:
                pairs.add((a, b))
        rounds += 1
    pairs = sorted(pairs)
    ea += [p[0] for p in pairs]
    eb += [p[1] for p in pairs]
    ea, eb = np.array(ea, dtype=np.int32), np.array(eb, dtype=np.int32)
    z = rel(truth[ea], truth[eb])
    z += np.column_stack([rng.normal(0, sigma_xy, (len(ea), 2)), rng.normal(0, sigma_th, len(ea))])
    z[:, 2] = wrap(z[:, 2])
    # covariance in the frame of pose a: R(-th_a) diag R(-th_a)^T  (Mapper.h:183-186)
    base = np.diag([sigma_xy ** 2, sigma_xy ** 2, sigma_th ** 2])
    cov = np.empty((len(ea), 3, 3))
    for k in range(len(ea)):
        t = -truth[ea[k], 2]
        R = np.array([[math.cos(t), -math.sin(t), 0], [math.sin(t), math.cos(t), 0], [0, 0, 1]])
        cov[k] = R @ base @ R.T
    # dead-reckoned initial guess from the odometry edges
    init = np.zeros_like(truth)
    init[0] = truth[0]
    for i in range(1, n_nodes):
        c, s = math.cos(init[i - 1, 2]), math.sin(init[i - 1, 2])
        init[i, 0] = init[i - 1, 0] + c * z[i - 1, 0] - s * z[i - 1, 1]
        init[i, 1] = init[i - 1, 1] + s * z[i - 1, 0] + c * z[i - 1, 1]
        init[i, 2] = wrap(init[i - 1, 2] + z[i - 1, 2])
    ids = np.arange(n_nodes, dtype=np.int32)
    return dict(ids=ids, init=init, truth=truth, edge_a=ea, edge_b=eb, z=z, cov=cov)
