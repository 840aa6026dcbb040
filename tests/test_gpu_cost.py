"""GPU parity: CostGreedyEndpoint::Cost and the matchers' tail (normalised cost + covariance) through
the C ABI vs the reference's own code -- bit-identical, the sum being taken in beam order."""
import os

import numpy as np
import pytest

from my_lidar_graph_slam_b200 import capi, synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _same(a, b):
    return np.array_equal(np.asarray(a, dtype=np.float64).view(np.int64), np.asarray(b, dtype=np.float64).view(np.int64))


def test_cost_golden_vectors(ctx):
    """tests/golden/scene_cost.npz was written by the unmodified reference objects."""
    from oracle import portapi as P
    g, c = np.load(os.path.join(GOLD, "scene_rooms.npz")), np.load(os.path.join(GOLD, "scene_cost.npz"))
    angles, traj, scans = g["angles"], g["traj"], g["scans"]
    assert np.array_equal(g["latest_geom"][:2].astype(int), g["latest_dense"].shape[::-1])
    nx, ny, mx, my, res = g["latest_geom"][:5]
    grid = capi.Grid.from_dense(ctx, g["latest_dense"], mx, my, res, apron=4)
    ks = sorted(set(c["scan"].tolist()))
    batch = capi.Scans([angles] * len(ks), [scans[k] for k in ks], np.zeros((len(ks), 3)),
                       range_min=0.02, range_max=30.0)
    pose_scan = [ks.index(k) for k in c["scan"]]
    for j, cs in enumerate(c["cost_sets"]):
        got, _ = capi.cost_greedy_endpoint(ctx, grid, batch, c["poses"], pose_scan, cost=tuple(cs))
        assert _same(got, c["cost"][:, j])
    # the tail: one best pose per scan of the batch
    first = [pose_scan.index(i) for i in range(len(ks))]
    for j, cs in enumerate(c["cost_sets"]):
        nc, cov, _ = capi.cost_tail(ctx, grid, batch, c["poses"][first], cost=tuple(cs))
        assert _same(nc, c["normalized"][first, j])
        assert np.array_equal(cov, c["cov"][first, j])
    del P


def test_cost_random_maps_ragged_scans_against_the_oracle(ctx):
    from oracle import backend
    R = backend()
    rng = np.random.default_rng(41)
    dense = np.where(rng.random((128, 192)) < 0.5, np.round(rng.uniform(1e-3, 0.999, (128, 192)), 2), 0.0)
    rm = R.RefMap.from_dense(dense, -3.0, -2.0)
    grid = capi.Grid.from_dense(ctx, dense, -3.0, -2.0, 0.05, apron=1)
    beams = [1, 7, 91, 360, 1081, 4500]                         # 4500 > one shared-memory round
    angles = [synth.beam_angles(n, 360.0) if n > 1 else np.array([0.3]) for n in beams]
    ranges = [rng.uniform(0.0, 4.0, n) for n in beams]
    smin, smax = rng.uniform(0.0, 0.3, len(beams)), rng.uniform(3.0, 6.0, len(beams))
    batch = capi.Scans(angles, ranges, np.zeros((len(beams), 3)), range_min=smin, range_max=smax)
    pose_scan = rng.integers(0, len(beams), 60)
    poses = np.stack([rng.uniform(-3.5, 7.0, 60), rng.uniform(-2.5, 5.0, 60), rng.uniform(-4, 4, 60)], axis=1)
    for K in (0, 1, 2, 3, 7):
        cs = (float(rng.uniform(0.0, 0.5)), float(rng.uniform(2.0, 5.0)), float(rng.uniform(0.02, 0.2)),
              float(rng.uniform(0.05, 0.6)), float(K), float(rng.uniform(0.1, 3.0)), float(rng.uniform(0.02, 1.0)))
        got, _ = capi.cost_greedy_endpoint(ctx, grid, batch, poses, pose_scan, cost=cs)
        want = [R.cost_greedy_endpoint(rm, poses[p], angles[s], ranges[s], scan_min_range=float(smin[s]),
                                       scan_max_range=float(smax[s]), cost=cs) for p, s in enumerate(pose_scan)]
        assert _same(got, want)
        assert len(set(want)) > 30
        nc, cov, _ = capi.cost_tail(ctx, grid, batch, poses[:len(beams)], cost=cs)
        for s in range(len(beams)):
            wn, _, wc = R.host_tail(rm, poses[s], angles[s], ranges[s], scan_min_range=float(smin[s]),
                                    scan_max_range=float(smax[s]), cost=cs)
            assert nc[s] == wn and np.array_equal(cov[s], wc)


def test_cost_host_fixups_leave_results_identical(ctx):
    """With a huge guard band nearly every beam takes the host (glibc) re-derivation path."""
    from scenes import room_scene
    from oracle import backend
    R = backend()
    world, angles, traj, builder = room_scene(seed=1)
    refmap = builder.latest_map()
    nx, ny, mx, my, res = refmap.geometry()
    grid = capi.Grid.from_dense(ctx, refmap.dense(), mx, my, res, apron=2)
    scan = synth.make_scan(world, traj[11], angles, np.random.default_rng(5))
    batch = capi.Scans([angles], [scan], [traj[11]], range_min=0.02, range_max=30.0)
    base_nc, base_cov, fix0 = capi.cost_tail(ctx, grid, batch, [traj[11]])
    wn, _, wc = R.host_tail(refmap, traj[11], angles, scan)
    assert base_nc[0] == wn and np.array_equal(base_cov[0], wc) and wn < 0.0
    try:
        ctx.set_edge_eps(0.3)
        nc, cov, fix = capi.cost_tail(ctx, grid, batch, [traj[11]])
    finally:
        ctx.set_edge_eps(0.0)          # restores the default
    assert fix > 1000 and fix0 < 5
    assert nc[0] == wn and np.array_equal(cov[0], wc)


def test_cost_edge_cases(ctx):
    dense = np.zeros((64, 64))
    dense[30:34, 30:34] = 0.7
    grid = capi.Grid.from_dense(ctx, dense, 0.0, 0.0, 0.05, apron=1)
    empty = capi.Scans([np.zeros(0)], [np.zeros(0)], np.zeros((1, 3)))
    got, fix = capi.cost_greedy_endpoint(ctx, grid, empty, [[1.0, 1.0, 0.0]])
    assert got.tolist() == [0.0] and fix == 0
    got, _ = capi.cost_greedy_endpoint(ctx, grid, empty, np.zeros((0, 3)), pose_scan=[])
    assert len(got) == 0
    one = capi.Scans([np.array([0.0])], [np.array([0.5])], np.zeros((1, 3)))
    with pytest.raises(capi.LgsError):
        capi.cost_greedy_endpoint(ctx, grid, one, [[1.0, 1.0, 0.0]], cost=(0.01, 20.0, 0.075, 0.1, 8, 1.0, 0.05))
    with pytest.raises(capi.LgsError):
        capi.cost_greedy_endpoint(ctx, grid, one, [[1.0, 1.0, 0.0]], pose_scan=[3])
