"""GPU parity across devices: loop detection sharded over 2+ GPUs of one box
must return exactly what one GPU returns, and what the reference's ScanMatcherBranchBound returns.
Skipped on a box with a single GPU."""
import os
import subprocess
import sys

import numpy as np
import pytest

from my_lidar_graph_slam_b200 import capi, synth

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BB = dict(node_height_max=6, range_x=2.0, range_y=2.0, range_theta=1.0, scan_range_max=20.0,
          score_range_min=0.01, score_range_max=20.0)


def _n_devices():
    return capi.device_count()


def _scene(n_submaps):
    world = synth.RoomsWorld(40.0, 5.0, seed=6)
    angles = synth.beam_angles(1081, 270.0)
    anchor = synth.trajectory(world, 1, seed=31)[0]
    maps = []
    for g in range(n_submaps):
        rng = np.random.default_rng(500 + g)
        start = None
        if g % 3 == 0:                                   # every third submap contains the query location
            start = (anchor[0] + rng.uniform(-0.4, 0.4), anchor[1] + rng.uniform(-0.4, 0.4), anchor[2])
            if not world.is_free(start[0], start[1], 0.6):
                start = tuple(anchor)
        traj = synth.trajectory(world, 8, step=0.3, seed=g, start=start) if start else \
            synth.trajectory(world, 8, step=0.3, seed=g)
        maps.append((traj, [synth.make_scan(world, p, angles, rng) for p in traj]))
    qrng = np.random.default_rng(9)
    scans, inits = [], []
    for k in range(3):
        t = anchor + np.array([0.15 * k, -0.1 * k, 0.03 * k])
        scans.append(synth.make_scan(world, t, angles, qrng))
        inits.append(t + np.array([0.3, -0.25, 0.08]))
    return angles, maps, scans, inits


def _build(ctx, angles, traj, scans):
    import bench
    grid, _ = bench.build_map_on_gpu(ctx, traj, angles, scans, apron=1)
    return grid, capi.Pyramid(ctx, grid, 6)


@pytest.mark.skipif(_n_devices() < 2, reason="needs at least two GPUs")
def test_group_detect_matches_one_gpu_and_the_reference(ctx):
    from oracle import backend
    R = backend()
    n_dev = min(_n_devices(), 8)
    n_sub = 13                                            # not a multiple of the group size: ragged shares
    angles, maps, qscans, qinits = _scene(n_sub)
    # one GPU: every submap on device 0
    single = [_build(ctx, angles, *m) for m in maps]
    scans = capi.Scans([angles] * len(qscans), qscans, qinits, range_min=0.02, range_max=30.0)
    pair_scan = np.repeat(np.arange(len(qscans), dtype=np.int32), n_sub)
    batch = capi.BbBatch(ctx, **BB)
    batch.upload_pairs(scans, pair_scan, [p for _, p in single] * len(qscans), 0.55)
    batch.run()
    want = batch.results_array()
    want_rec = batch.records()
    assert 2 <= int((want["found"] != 0).sum()) < len(want)
    # the group: submap i on member i % G, all members searched concurrently from one process
    group = capi.Group(list(range(n_dev)))
    sharded = [_build(group.ctxs[g % n_dev], angles, *m) for g, m in enumerate(maps)]
    det = capi.GroupBb(group, **BB)
    pyr = [p for _, p in sharded] * len(qscans)
    for rep in range(3):                                  # later repetitions reuse pools / gather buffer
        got = det.detect(scans, pair_scan, pyr, 0.55)
        for f in ("found", "ix", "iy", "it", "win_x", "win_y", "win_t", "step_t", "score"):
            assert np.array_equal(got[f], want[f]), (rep, f)
        rec = det.records()
        assert rec.tobytes() == want_rec.tobytes()        # the exchanged 32-byte records, pair order
    # ... and the reference's CPU matcher on a few pairs
    for q in (0, 3, n_sub + 6, 2 * n_sub + 12):
        k, g = divmod(q, n_sub)
        dense = single[g][0].download()
        refmap = R.RefMap.from_dense(dense, single[g][0].min_x, single[g][0].min_y)
        ref = R.bb_match(refmap, angles, qscans[k], qinits[k], pyramid=refmap.pyramid(6), thr=0.55,
                         height_max=6, range_x=2.0, range_y=2.0, range_theta=1.0, scan_range_max=20.0,
                         score_range_min=0.01, score_range_max=20.0)
        assert (int(got[q]["found"]), int(got[q]["ix"]), int(got[q]["iy"]), int(got[q]["it"])) == \
            (ref.found, ref.ix, ref.iy, ref.it)
        if ref.found:
            assert float(got[q]["score"]) == ref.score
    det.close()
    for g_, p_ in sharded:
        p_.close()
        g_.close()
    group.close()


@pytest.mark.skipif(_n_devices() < 2, reason="needs at least two GPUs")
def test_comm_all_gather_of_records_between_processes():
    """One process per GPU (the bench's launch shape): every rank's kernel writes its records into its own
    slice of the receive buffer and an in-place NCCL all-gather on the context stream completes it."""
    n = min(_n_devices(), 4)
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", "29517",
                        os.path.join(ROOT, "tools", "comm_check.py")],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:]
    assert f"COMM_CHECK_OK world={n}" in p.stdout, p.stdout[-3000:]
