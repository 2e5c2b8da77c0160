"""Generates tests/golden/occupancy_golden.npz from the UNMODIFIED reference (oracle/_ref/libkarto_ref.so):
karto::OccupancyGrid::CreateFromScans (Karto.h:5946-5961) on small synthetic mapping runs.  Run in the
build container:
    python tests/golden/make_occupancy_golden.py
Each case stores the inputs (ranges, sensor poses, parameters) and the reference's grid: dimensions, offset,
the cell bytes and digests + sums of the pass / hit counters."""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import helpers as H  # noqa: E402
from slam_toolbox_b200 import synth  # noqa: E402
from oracle import karto_ref as R  # noqa: E402


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


CASES = [
    # name, seed, scans, resolution, range threshold, min_pass_through, occupancy_threshold  (negative = reference defaults)
    ("default", 51, 16, 0.05, 12.0, -1, -1.0),
    ("coarse_strict", 52, 12, 0.1, 12.0, 5, 0.3),
    ("short_threshold", 53, 10, 0.05, 4.0, 0, 0.05),
    ("one_scan", 54, 1, 0.05, 12.0, -1, -1.0),
]


def main():
    out = {}
    for name, seed, n, res, rt, mp, th in CASES:
        run = synth.make_mapping_run(seed, n)
        R.init_laser(**H.LASER)
        R.lib().kref_laser_set_range_threshold(rt)
        scans = H.ref_scans(run["ranges"], run["poses"])
        g = R.occupancy(scans, res, mp, th)
        R.lib().kref_laser_set_range_threshold(H.LASER["range_threshold"])
        out[f"{name}/ranges"] = run["ranges"]
        out[f"{name}/poses"] = run["poses"]
        out[f"{name}/params"] = np.array([res, rt, mp, th])
        out[f"{name}/dims"] = np.array([g["width"], g["height"], g["stride"]])
        out[f"{name}/offset"] = g["offset"]
        out[f"{name}/cells"] = g["cells"]
        out[f"{name}/pass_sha"] = np.array([digest(g["passes"])])
        out[f"{name}/hits_sha"] = np.array([digest(g["hits"])])
        out[f"{name}/sums"] = np.array([int(g["passes"].sum()), int(g["hits"].sum())])
        print(name, (g["width"], g["height"], g["stride"]), "occupied", int((g["cells"] == 100).sum()), "free",
              int((g["cells"] == 255).sum()), "pass", int(g["passes"].sum()), "hits", int(g["hits"].sum()))
    path = os.path.join(HERE, "occupancy_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
