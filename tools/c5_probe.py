import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from my_lidar_graph_slam_b200 import capi
ctx = capi.Context(0)
side = int(os.environ.get("C5_SIDE", "8000"))
rng = np.random.default_rng(1)
tile = np.where(rng.random((1000, 1000)) < 0.3, rng.random((1000, 1000)), 0.0)
dense = np.tile(tile, (side // 1000, side // 1000))
g = capi.Grid.from_dense(ctx, dense, 0.0, 0.0, 0.05, apron=1)
for rep in range(3):
    ctx.synchronize()
    t0 = time.perf_counter()
    ctx.timer_start()
    p = capi.Pyramid(ctx, g, 6)
    ms = ctx.timer_stop()
    wall = time.perf_counter() - t0
    print(f"pyramid {side}x{side} x 7 levels: events {ms:.3f} ms, wall {wall * 1e3:.3f} ms, "
          f"{side * side * 6 * 16 / ms / 1e6:.1f} GB/s algorithmic (16 B x cells x 6 built levels)")
    p.close()
