"""cfg3: offline synchronous replay of N synthetic posed scans through the reference's karto::Mapper::Process,
reference CPU matcher vs GPU matcher (link-time seam), both with the GPU ScanSolver adapter.
    python tools/replay_bench.py [n_scans]
Prints one JSON line with scans/s for both, match/solve counts and whether the final poses are identical."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "integration"))
import replay  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
ranges, odom, truth = replay.make_trajectory(6, n)
params = dict(replay.YAML_PARAMS, loop_search_space_dimension=4.0)
out = {"workload": f"cfg3 replay: {n} posed 1081-beam scans, slam_toolbox YAML parameters (smear 0.1 -> order-dependent raster), loop search 4 m"}
res = {}
for which in ("b200", "ref"):
    r = replay.run(which, ranges, odom, params, map_resolution=0.05 if which == "b200" else 0.0)
    res[which] = r
    out[which] = {"scans_kept": int(r["scans"]), "edges": int(r["edges"]), "process_s": float(r["process_seconds"]),
                  "scans_per_s": float(r["scans"]) / float(r["process_seconds"]), "solver_computes": int(r["solver_computes"]),
                  "solver_ms": float(r["solver_ms"]), "match_calls": int(r["match_calls"])}
b = res["b200"]
out["map_publish"] = {"scans": int(b["scans"]), "resolution": 0.05, "grid": [int(v) for v in b["map_gpu_dims"][:2]],
                      "reference_cpu_ms": 1e3 * float(b["map_cpu_seconds"]), "b200_ms_host_to_host": 1e3 * float(b["map_gpu_seconds"]),
                      "identical_cells": bool(np.array_equal(b["map_cpu_cells"], b["map_gpu_cells"]))}
out["identical_poses"] = bool(np.array_equal(res["ref"]["poses"], res["b200"]["poses"]))
out["speedup"] = out["b200"]["scans_per_s"] / out["ref"]["scans_per_s"]
print(json.dumps(out))
