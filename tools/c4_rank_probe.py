"""Probe: one rank's share of the 8-GPU batched loop detection (16 scans x 250 submaps) on one GPU, for
different sub-batch splits -- how much of the step is latency that only concurrent sub-batches hide."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from my_lidar_graph_slam_b200 import capi, synth  # noqa: E402

NS, NM, Q = int(os.environ.get("NS", 16)), int(os.environ.get("NM", 250)), 64
ctx = capi.Context(0)
world = synth.RoomsWorld(60.0, 5.0, seed=4)
angles = synth.beam_angles(1081, 270.0)
anchor = synth.trajectory(world, 1, seed=77)[0]
pyr = []
for g in range(0, 2 * NM, 2):
    traj, scans = bench.c4_submap_scans(world, angles, g, 8, anchor)
    grid, _ = bench.build_map_on_gpu(ctx, traj, angles, scans, apron=1)
    pyr.append(capi.Pyramid(ctx, grid, 6))
qrng = np.random.default_rng(5)
synth.make_scan(world, anchor, angles, qrng)
qs, qi = [], []
for k in range(Q):
    t = anchor + np.array([0.3 * np.cos(k), 0.3 * np.sin(k), 0.04 * k])
    qs.append(synth.make_scan(world, t, angles, qrng))
    qi.append(t + np.array([0.3, -0.2, 0.08]))
mine = [k for k in range(Q) if k % (Q // NS) == 0]
for sub in (16, 8):
    groups = []
    for k0 in range(0, len(mine), sub):
        ks = mine[k0:k0 + sub]
        groups.append(dict(scans=capi.Scans([angles] * len(ks), [qs[k] for k in ks], [qi[k] for k in ks],
                                            range_min=0.02, range_max=30.0),
                           pair=np.repeat(np.arange(len(ks), dtype=np.int32), NM), pyr=pyr * len(ks),
                           batch=capi.BbBatch(ctx, **bench.BB)))

    def step():
        for g in groups:
            g["batch"].upload_pairs(g["scans"], g["pair"], g["pyr"], 0.6)
            g["batch"].run()
        return [g["batch"].results_array() for g in groups]

    for _ in range(3):
        step()
    ctx.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        step()
    ctx.synchronize()
    e2e = (time.perf_counter() - t0) / 10
    for g in groups:
        g["batch"].upload_pairs(g["scans"], g["pair"], g["pyr"], 0.6)
    ctx.timer_start()
    for _ in range(10):
        for g in groups:
            g["batch"].run()
    dev = ctx.timer_stop() / 10
    print(f"{NS} scans x {NM} submaps in sub-batches of {sub} scans ({len(groups)} groups): kernels {dev:.2f} ms, "
          f"end to end {1e3 * e2e:.2f} ms per step")
    for g in groups:
        g["batch"].close()
for p in pyr:
    p.close()
ctx.close()
