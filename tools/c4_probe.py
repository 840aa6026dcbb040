import sys, json, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from my_lidar_graph_slam_b200 import capi
ctx = capi.Context(0)
n = int(os.environ.get("C4_SUBMAPS", "500"))
out = bench.run_c4(ctx, 0, 1, 0, lambda: None, lambda x: x, n, 10, False)
print(n, {k: out[k] for k in ("loop_queries_per_s", "ms_per_query_batch", "ms_per_query_batch_e2e", "nodes_per_level_rank0")}, out.get("batched"))
