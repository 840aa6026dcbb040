import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from my_lidar_graph_slam_b200 import capi
ctx = capi.Context(0)
out = bench.run_c3(ctx, int(os.environ.get("C3_DISTINCT", "256")), int(os.environ.get("C3_SCANS", "4096")), False)
print({k: out[k] for k in ("scans_per_s", "cell_updates_per_s", "cell_updates_per_scan")})
