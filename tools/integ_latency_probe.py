"""Probe: wall time of ONE lgs_grid_integrate_scans call for small batches (the per-frame builder case)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from my_lidar_graph_slam_b200 import capi, synth  # noqa: E402

beams = int(os.environ.get("BEAMS", 180))
world = synth.RoomsWorld(24.0, 4.0, seed=3)
angles = synth.beam_angles(beams, 180.0)
traj = synth.trajectory(world, 12, step=0.1, seed=3)
noise = np.random.default_rng(4)
hits = [capi.scan_hit_points(p, angles, synth.make_scan(world, p, angles, noise), 0.01, 20.0)[0] for p in traj]
ctx = capi.Context(0)
grid = capi.Grid(ctx, 640, 640, -16.0, -16.0, 0.05, apron=1)
for n in (1, 10):
    packed = capi.PackedHits(traj[:n, :2], hits[:n])
    for _ in range(5):
        capi.integrate_packed(ctx, grid, packed)
    t0 = time.perf_counter()
    reps = 200
    for _ in range(reps):
        capi.integrate_packed(ctx, grid, packed)
    print(f"{n} scan(s) of {beams} beams: {1e6 * (time.perf_counter() - t0) / reps:.0f} us per call")
