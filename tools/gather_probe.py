import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from my_lidar_graph_slam_b200 import capi
ctx = capi.Context(0)
for lanes in (32, 25):
    for aligned in (True, False):
        for local in (True, False):
            print(f"rows of {lanes} doubles, aligned={aligned}, local={local}: "
                  f"{capi.measure_gather_peak(ctx, 960, 640, lanes, aligned, local):9.1f} GB/s")
