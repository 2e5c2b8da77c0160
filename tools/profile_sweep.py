"""Small driver for ncu: uploads the cfg2 workload (1 query x N candidates) and runs the sweep a few times."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from slam_toolbox_b200 import api

n = int(sys.argv[1]) if len(sys.argv) > 1 else 296
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 2
qr, qp, cr, cp, cs = bench.make_inputs(0, n, 1, 1)
laser = api.LaserRangeFinder()
mapper = api.MapperParams(**{k: (bool(v) if k == "use_response_expansion" else v) for k, v in bench.LOOP_MAPPER.items()})
sm = api.ScanMatcher.Create(mapper, *bench.LOOP_GRID)
if len(sys.argv) > 3 and sys.argv[3] == "generic":
    sm.set_option("force_generic_sweep", 1)
sm.batch_upload(api.ScanBlock(qr, qp, laser), api.ScanBlock(cr, cp, laser), cs, None, False)
for _ in range(runs):
    sm.batch_run()
r = sm.batch_fetch()
print("kernel ms", sm.batch_kernel_ms(), "best", r[0].max())
