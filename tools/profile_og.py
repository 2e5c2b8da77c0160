"""Small driver for ncu / compute-sanitizer: one occupancy-grid build over N synthetic scans (default 1500)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from slam_toolbox_b200 import api, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
builds = int(sys.argv[2]) if len(sys.argv) > 2 else 2
run = synth.make_mapping_run(3, n, world=synth.make_world(3, size=60.0), odd_readings=False)
blk = api.ScanBlock(run["ranges"], run["poses"], api.LaserRangeFinder())
g = api.OccupancyGrid(0.05, blk.laser)
g.AddScans(blk)
for _ in range(builds):
    g.Build()
cells, ps, ht = g.GetData(counters=True)
print("build ms", g.kernel_ms(), "grid", g.GetWidth(), g.GetHeight(), "updates", int(ps.sum()), "hits", int(ht.sum()))
