"""Probe: lgs_gs_match at the launcher's LoopDetectorGridSearch defaults (2 m x 2 m at 0.05 m, 0.5 rad at
0.005 rad, 1081 beams) for GS_Q queries against one map."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from my_lidar_graph_slam_b200 import capi  # noqa: E402

Q = int(os.environ.get("GS_Q", 16))
traj, map_scans, angles, ranges, inits = bench.c2_workload(Q, seed=1)
ctx = capi.Context(0)
grid, _ = bench.build_map_on_gpu(ctx, traj, angles, map_scans, apron=32)
scans = capi.Scans([angles] * Q, ranges, inits, range_min=0.02, range_max=30.0)
capi.pin(ctx, scans.angles, scans.ranges)
capi.gs_match(ctx, scans, [grid] * Q)
for _ in range(3):
    t0 = time.perf_counter()
    out = capi.gs_match(ctx, scans, [grid] * Q)
    dt = time.perf_counter() - t0
    hyp = sum(r.n_scored for r in out)
    print(f"{Q} queries: {1e3 * dt:.2f} ms, {Q / dt:.0f} queries/s, {hyp / dt:.3e} hypotheses/s, "
          f"found {sum(r.found for r in out)}, fixups {sum(r.n_fixups for r in out)}")
