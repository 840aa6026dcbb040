"""One device-only branch-and-bound run for ncu (BB_SUBMAPS submaps x 1 scan): warm-up runs, then the captured one."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from my_lidar_graph_slam_b200 import capi, synth  # noqa: E402

NM = int(os.environ.get("BB_SUBMAPS", 500))
NS = int(os.environ.get("BB_SCANS", 1))
ctx = capi.Context(0)
for kv in [x for x in os.environ.get("BB_OPTS", "").split(",") if x]:
    k, v = kv.split("=")
    ctx.set_option(k, float(v))
world = synth.RoomsWorld(60.0, 5.0, seed=4)
angles = synth.beam_angles(1081, 270.0)
anchor = synth.trajectory(world, 1, seed=77)[0]
pyr = []
for g in range(NM):
    traj, scans = bench.c4_submap_scans(world, angles, g, 8, anchor)
    grid, _ = bench.build_map_on_gpu(ctx, traj, angles, scans, apron=1)
    pyr.append(capi.Pyramid(ctx, grid, 6))
qrng = np.random.default_rng(5)
qs, qi = [], []
for k in range(NS):
    t = anchor + np.array([0.2 + 0.3 * np.cos(k), -0.1 + 0.3 * np.sin(k), 0.05 + 0.04 * k])
    qs.append(synth.make_scan(world, t, angles, qrng))
    qi.append(t + np.array([0.4, -0.3, 0.1]))
scans = capi.Scans([angles] * NS, qs, qi, range_min=0.02, range_max=30.0)
batch = capi.BbBatch(ctx, **bench.BB)
batch.upload_pairs(scans, np.repeat(np.arange(NS, dtype=np.int32), NM), pyr * NS, 0.6)
for _ in range(3):
    batch.run()
    res = batch.results_array()
print("found", int((res["found"] != 0).sum()), "path", batch.path())
ctx.close()
