"""Probe: branch-and-bound loop detection (C4 shape) on one GPU -- device time per run, per-phase split
of the persistent kernel, and the effect of its build / residency / cost-model options.

  BB_SUBMAPS (500)  submaps          BB_SCANS (1)  query scans per batch
  BB_BLOCKS ("0")  CTAs per SM to try (0 = occupancy limit)      BB_REPS (20)
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from my_lidar_graph_slam_b200 import capi, synth  # noqa: E402

NM = int(os.environ.get("BB_SUBMAPS", 500))
NS = int(os.environ.get("BB_SCANS", 1))
REPS = int(os.environ.get("BB_REPS", 20))
BLOCKS = [int(v) for v in os.environ.get("BB_BLOCKS", "0").split(",")]
EXTRA = os.environ.get("BB_OPTS", "")      # e.g. "bb_cost_g1=50,bb_cost_g4=20"

ctx = capi.Context(0)
world = synth.RoomsWorld(60.0, 5.0, seed=4)
angles = synth.beam_angles(1081, 270.0)
anchor = synth.trajectory(world, 1, seed=77)[0]
pyr = []
t0 = time.perf_counter()
for g in range(NM):
    traj, scans = bench.c4_submap_scans(world, angles, g, 8, anchor)
    grid, _ = bench.build_map_on_gpu(ctx, traj, angles, scans, apron=1)
    pyr.append(capi.Pyramid(ctx, grid, 6))
print(f"built {NM} submaps in {time.perf_counter() - t0:.1f} s", flush=True)
qrng = np.random.default_rng(5)
qs, qi = [], []
for k in range(NS):
    t = anchor + np.array([0.2 + 0.3 * np.cos(k), -0.1 + 0.3 * np.sin(k), 0.05 + 0.04 * k])
    qs.append(synth.make_scan(world, t, angles, qrng))
    qi.append(t + np.array([0.4, -0.3, 0.1]))
scans = capi.Scans([angles] * NS, qs, qi, range_min=0.02, range_max=30.0)
pair = np.repeat(np.arange(NS, dtype=np.int32), NM)
plist = pyr * NS


def measure(label):
    batch = capi.BbBatch(ctx, **bench.BB)
    batch.upload_pairs(scans, pair, plist, 0.6)
    for _ in range(3):
        batch.run()
        res = batch.results_array()
    ctx.synchronize()
    t0 = time.perf_counter()
    for _ in range(REPS):
        batch.upload_pairs(scans, pair, plist, 0.6)
        batch.run()
        res = batch.results_array()
    e2e = (time.perf_counter() - t0) / REPS
    batch.upload_pairs(scans, pair, plist, 0.6)
    ctx.synchronize()
    ctx.timer_start()
    for _ in range(REPS):
        batch.run()
    dev = ctx.timer_stop() / REPS
    levels, gathers = batch.work()
    line = (f"{label}: kernels {dev * 1e3:.1f} us, e2e {e2e * 1e6:.1f} us per run; path {batch.path()}; levels {levels}; "
            f"gathers issued {1.0 - batch.skipped_gathers() / max(gathers, 1):.3f} of nodes x beams")
    try:
        ctx.set_option("bb_host_timing", 1)
        batch.run()
        us, g = batch.phase_times()
        line += "\n    phases us " + " ".join(f"{u:.1f}" for u in us) + " | lanes/node " + " ".join(str(int(x)) for x in g)
    except capi.LgsError:
        pass
    finally:
        ctx.set_option("bb_host_timing", 0)
    print(line, flush=True)
    found = int((res["found"] != 0).sum())
    batch.close()
    return found, res


ctx.set_option("bb_sync", 1)
found_ref, res_ref = measure("exact path (level-synchronous)")
ctx.set_option("bb_sync", 0)
for kv in [x for x in EXTRA.split(",") if x]:
    k, v = kv.split("=")
    ctx.set_option(k, float(v))
for bps in BLOCKS:
    ctx.set_option("bb_blocks_per_sm", bps)
    if os.environ.get("BB_BOTH"):
        ctx.set_option("bb_early_reject", 0)
        measure(f"device-only, no early rejection, CTAs/SM {bps or 'max'}")
        ctx.set_option("bb_early_reject", 1)
    found, res = measure(f"device-only, CTAs/SM {bps or 'max'}")
    same = all(np.array_equal(res[f], res_ref[f]) for f in ("found", "ix", "iy", "it", "score"))
    print(f"    found {found} (exact path {found_ref}); identical to the exact path: {same}", flush=True)
ctx.close()
