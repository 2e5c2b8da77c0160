"""Geometry / plan explorer for the sweep kernels: for each (search dimension, range threshold) and kernel variant,
uploads 1 query x N candidates, prints the plan, the kernel time and pairs/s, and checks that all variants agree.
usage: tile_sweep.py [N=296] [geoms=4:12,8:12,4:20,8:20] [variants=fast,tile:1,tile:2,tile:4,tile:8,generic]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from slam_toolbox_b200 import api

n = int(sys.argv[1]) if len(sys.argv) > 1 else 296
geoms = sys.argv[2] if len(sys.argv) > 2 else "4:12,8:12,4:20,8:20"
variants = (sys.argv[3] if len(sys.argv) > 3 else "fast,tile:1,tile:2,tile:4,tile:8,generic").split(",")
chain = int(sys.argv[4]) if len(sys.argv) > 4 else 1
qr, qp, cr, cp, cs = bench.make_inputs(0, n, chain, 1)
laser = api.LaserRangeFinder()
mapper = api.MapperParams(**{k: (bool(v) if k == "use_response_expansion" else v) for k, v in bench.LOOP_MAPPER.items()})
for gstr in geoms.split(","):
    dim, rt = (float(v) for v in gstr.split(":"))
    sm = api.ScanMatcher.Create(mapper, dim, 0.05, 0.03, rt)
    q, c = api.ScanBlock(qr, qp, laser), api.ScanBlock(cr, cp, laser)
    ref = None
    for v in variants:
        sm.set_option("force_generic_sweep", 1 if v == "generic" else 0)
        kind, _, arg = v.partition(":")
        sm.set_option("sweep_kernel", {"fast": 1, "tile": 2, "generic": 0}[kind])
        cl, _, ch = arg.partition(":")
        sm.set_option("sweep_cluster", int(cl) if cl else 0)
        sm.set_option("sweep_chunks", int(ch) if ch else 0)
        sm.batch_upload(q, c, cs, None, False)
        info, plan = sm.batch_info(), sm.batch_tile_info()
        if kind != "generic" and info["kernel"] != kind:
            print(json.dumps({"geom": gstr, "variant": v, "skipped": info["kernel"], "refused": info["refused_reason"], "tile_refused": plan["refused_reason"]}))
            continue
        ms = []
        for _ in range(3):
            sm.batch_run()
            ms.append(sm.batch_kernel_ms())
        out = sm.batch_fetch()
        best = sm.batch_best()
        same = True
        if ref is None:
            ref = (out, best)
        else:
            same = all(np.array_equal(a, b) for a, b in zip(out, ref[0])) and all(np.array_equal(a, b) for a, b in zip(best, ref[1]))
        print(json.dumps({"geom": gstr, "variant": v, "kernel": info["kernel"], "ctas": info["ctas"], "plan": plan if kind == "tile" else None,
                          "kernel_ms": round(min(ms), 4), "pairs_per_s": round(n / (min(ms) * 1e-3)), "same_as_first": bool(same),
                          "edge": info["edge_beams"], "far": info["far_beams"], "best": float(out[0].max())}), flush=True)
    sm.close()
