"""Probe: lgs_cost_tail on a C2-like batch (M matches x 1081 beams); prints wall time per call."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from my_lidar_graph_slam_b200 import capi  # noqa: E402

M = int(os.environ.get("TAIL_M", 1000))
traj, map_scans, angles, ranges, inits = bench.c2_workload(M, seed=1)
ctx = capi.Context(0)
grid, _ = bench.build_map_on_gpu(ctx, traj, angles, map_scans, apron=32)
scans = capi.Scans([angles] * M, ranges, inits, range_min=0.02, range_max=30.0)
best = np.asarray(inits, dtype=np.float64)
if os.environ.get("TAIL_PIN", "1") == "1":
    capi.pin(ctx, scans.angles, scans.ranges)
capi.cost_tail(ctx, grid, scans, best)
for _ in range(3):
    t0 = time.perf_counter()
    nc, cov, fix = capi.cost_tail(ctx, grid, scans, best)
    dt = time.perf_counter() - t0
    print(f"{M} tails: {1e3 * dt:.3f} ms per call, {M / dt:.0f} tails/s, fixups {fix}, mean cost {nc.mean():.4f}")
one = capi.Scans([angles], [ranges[0]], [inits[0]], range_min=0.02, range_max=30.0)
capi.cost_tail(ctx, grid, one, best[:1])
t0 = time.perf_counter()
for _ in range(100):
    capi.cost_tail(ctx, grid, one, best[:1])
print(f"single tail: {1e4 * (time.perf_counter() - t0):.1f} us per call")
