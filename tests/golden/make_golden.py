"""Generates tests/golden/matcher_golden.npz from the UNMODIFIED reference (oracle/_ref/libkarto_ref.so,
built by oracle/Makefile from /root/reference).  Run in the build container:
    python tests/golden/make_golden.py
Each case stores the inputs (ranges, poses) and what the reference's ScanMatcher::MatchScan returned
(response, mean, covariance) plus digests of its correlation grid, lookup table and the integer
response volume of the coarse pass.  The fixtures pin the C-port oracle and the CUDA path on machines
where /root/reference does not exist (the GPU box)."""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import helpers as H  # noqa: E402
from slam_toolbox_b200 import synth  # noqa: E402


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def volume_from(grid, offsets, xs_idx):
    """Integer response volume from the reference's grid bytes + lookup table (Mapper.cpp:1190-1201)."""
    data, size = grid["data"].astype(np.int64), grid["data_size"]
    vol = np.zeros((len(xs_idx), offsets.shape[0]), dtype=np.int32)
    for a in range(offsets.shape[0]):
        off = offsets[a].astype(np.int64)
        valid = off != 2147483647
        idx = xs_idx[:, None] + off[None, valid]
        ok = (idx >= 0) & (idx < size)
        vol[:, a] = np.where(ok, data[np.clip(idx, 0, size - 1)], 0).sum(axis=1)
    return vol


CASES = [
    # name, mapper, grid, seed, buffer_len, inf_frac, nan_frac, penalize, refine
    ("seq_k03", H.MAPPER_SEQ, H.GRID_SEQ, 11, 6, 0.0, 0.0, True, True),
    ("seq_yaml_inf", H.MAPPER_SEQ, H.GRID_SEQ_YAML, 12, 4, 0.04, 0.01, True, True),
    ("loop_chain5", H.MAPPER_LOOP, H.GRID_LOOP, 13, 5, 0.02, 0.0, False, False),
    ("loop_refine", H.MAPPER_LOOP, H.GRID_LOOP, 14, 3, 0.0, 0.0, False, True),
    ("small", H.MAPPER_LOOP, H.GRID_SMALL, 15, 2, 0.0, 0.0, True, True),
]


def main():
    out = {}
    for name, mapper, grid, seed, blen, inf_frac, nan_frac, pen, refine in CASES:
        case = synth.make_sequential_case(seed, buffer_len=blen, inf_frac=inf_frac, nan_frac=nan_frac)
        rm = H.ref_matcher(mapper, grid)
        base = H.ref_scans(case["base_ranges"], case["base_poses"])
        q = H.ref_scans(case["query_ranges"], case["query_pose"], 1000)[0]
        resp, mean, cov = rm.match(q, base, pen, refine)
        rm.raster(q, base)
        g = rm.grid()
        off = rm.offsets(q, case["query_pose"][2], mapper["coarse_search_angle_offset"], mapper["coarse_angle_resolution"])
        so, sr = H.coarse_search(grid)
        # pose -> grid index exactly like ScanMatcher::operator() (Mapper.cpp:660-662)
        nx = int(np.floor(so[0] * 2.0 / sr[0] + 0.5)) + 1
        xs = np.array([-so[0] + k * sr[0] for k in range(nx)])
        gx = np.floor((case["query_pose"][0] + xs - g["offset"][0]) * g["scale"] + 0.5).astype(np.int64) + g["roi"][0]
        gy = np.floor((case["query_pose"][1] + xs - g["offset"][1]) * g["scale"] + 0.5).astype(np.int64) + g["roi"][1]
        pos = (gy[:, None] * g["stride"] + gx[None, :]).reshape(-1)
        vol = volume_from(g, off, pos).reshape(nx, nx, off.shape[0])
        # the volume must reproduce the reference's own coarse response (unpenalised)
        r2, _, _ = rm.correlate(q, case["query_pose"], so, sr, mapper["coarse_search_angle_offset"], mapper["coarse_angle_resolution"], False, False)
        assert r2 == min(1.0, vol.max() / (len(case["query_ranges"]) * 100.0)), (name, r2, vol.max())
        out[f"{name}/base_ranges"] = case["base_ranges"]
        out[f"{name}/base_poses"] = case["base_poses"]
        out[f"{name}/query_ranges"] = case["query_ranges"]
        out[f"{name}/query_pose"] = case["query_pose"]
        out[f"{name}/flags"] = np.array([int(pen), int(refine)])
        out[f"{name}/response"] = np.array([resp])
        out[f"{name}/mean"] = mean
        out[f"{name}/cov"] = cov
        out[f"{name}/grid_sha"] = np.array([digest(g["data"])])
        out[f"{name}/grid_nonzero"] = np.array([int((g["data"] > 0).sum()), int(g["data"].astype(np.int64).sum())])
        out[f"{name}/offsets_sha"] = np.array([digest(off)])
        out[f"{name}/volume_sha"] = np.array([digest(vol)])
        out[f"{name}/volume_argmax"] = np.array([int(vol.argmax()), int(vol.max())])
        out[f"{name}/kernel"] = rm.kernel()
        print(name, "response", resp, "argmax", int(vol.argmax()), int(vol.max()), "grid nonzero", int((g["data"] > 0).sum()))
    np.savez_compressed(os.path.join(HERE, "matcher_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "matcher_golden.npz"), os.path.getsize(os.path.join(HERE, "matcher_golden.npz")), "bytes")


if __name__ == "__main__":
    main()
