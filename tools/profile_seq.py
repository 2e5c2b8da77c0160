"""Small driver for ncu: a few cfg1 sequential matches (k_stamp + k_correlate, coarse + fine) through b200sm_match."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from slam_toolbox_b200 import api, synth
smear = float(sys.argv[1]) if len(sys.argv) > 1 else 0.03
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
mapper = api.MapperParams(**{k: (bool(v) if k == "use_response_expansion" else v)
                             for k, v in dict(bench.LOOP_MAPPER, coarse_search_angle_offset=math.radians(5.0)).items()})
sm = api.ScanMatcher.Create(mapper, 1.0, 0.01, smear, 12.0)
laser = api.LaserRangeFinder()
c = synth.make_sequential_case(100, buffer_len=10)
q, b = api.ScanBlock(c["query_ranges"][None, :], c["query_pose"][None, :], laser), api.ScanBlock(c["base_ranges"], c["base_poses"], laser)
for _ in range(n):
    r = sm.MatchScan(q, b, True, True)
print("response", r[0])
