"""GPU parity: sliding-window-max maps vs the reference's PrecomputeGridMap(s) (cell values bit-exact)."""
import numpy as np
import pytest

from my_lidar_graph_slam_b200 import capi

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a).view(np.int64)


@pytest.mark.parametrize("win", [1, 2, 3, 5, 8, 64, 100])
def test_precompute_matches_reference(ctx, win):
    from oracle import backend
    R = backend()
    from scenes import room_scene
    _, _, _, builder = room_scene(seed=1)
    refmap = builder.latest_map()
    nx, ny, mx, my, res = refmap.geometry()
    grid = capi.Grid.from_dense(ctx, refmap.dense(), mx, my, res, apron=8)
    got = grid.precompute(win).download()
    want = refmap.precompute(win).dense()
    assert np.array_equal(_bits(got), _bits(want))


@pytest.mark.parametrize("shape", [(64, 64), (64, 192), (128, 64)])
def test_pyramid_matches_reference_small_maps(ctx, shape):
    """Windows up to 2^7 = 128 exceed one or both map extents (n < 2w and n < w branches)."""
    from oracle import backend
    R = backend()
    rng = np.random.default_rng(11)
    dense = np.where(rng.random(shape) < 0.3, rng.uniform(0.001, 0.999, shape), 0.0)
    refmap = R.RefMap.from_dense(dense, -3.2, 1.6)
    grid = capi.Grid.from_dense(ctx, refmap.dense(), -3.2, 1.6, 0.05, apron=4)
    pyr = capi.Pyramid(ctx, grid, 7)
    ref = refmap.pyramid(7)
    for h in range(8):
        assert np.array_equal(_bits(pyr.download(h)), _bits(ref[h].dense())), f"level {h}"


def test_pyramid_matches_reference_scene(ctx):
    from scenes import room_scene
    _, _, _, builder = room_scene(seed=2)
    refmap = builder.local_map(0)
    nx, ny, mx, my, res = refmap.geometry()
    grid = capi.Grid.from_dense(ctx, refmap.dense(), mx, my, res, apron=4)
    pyr = capi.Pyramid(ctx, grid, 6)
    ref = refmap.pyramid(6)
    for h in range(7):
        assert np.array_equal(_bits(pyr.download(h)), _bits(ref[h].dense())), f"level {h}"


def test_precompute_properties_at_full_size(ctx):
    """8000 x 8000 (config C5): idempotence of window 1, monotonicity across levels, and the
    closed form checked on sampled cells (the oracle would take minutes at this size)."""
    rng = np.random.default_rng(5)
    n = 8000
    dense = np.zeros((n, n))
    ys, xs = rng.integers(0, n, 400_000), rng.integers(0, n, 400_000)
    dense[ys, xs] = rng.uniform(0.001, 0.999, ys.shape)
    grid = capi.Grid.from_dense(ctx, dense, 0.0, 0.0, 0.05, apron=2)
    pyr = capi.Pyramid(ctx, grid, 6)
    prev = pyr.download(0)
    assert np.array_equal(_bits(prev), _bits(dense))
    for h in range(1, 7):
        cur = pyr.download(h)
        assert (cur >= prev).all()
        w = 1 << h
        for _ in range(200):
            y, x = int(rng.integers(0, n)), int(rng.integers(0, n))
            y0, x0 = min(y, n - w), min(x, n - w)
            assert cur[y, x] == dense[y0:y0 + w, x0:x0 + w].max()
        prev = cur


def test_pyramid_apron_is_zero_when_the_pool_hands_back_dirty_memory(ctx):
    """Pyramid slabs come from the stream-ordered pool and only their apron is cleared.  Dirty the
    pool with an all-ones pyramid of the same size, free it, build the real one and match a scan
    whose window hangs over the map edge: stale apron cells would change the scores."""
    from my_lidar_graph_slam_b200 import synth
    from oracle import backend
    R = backend()
    world = synth.World(16.0, 16.0, 5, seed=4)
    angles = synth.beam_angles(361, 180.0)
    traj = synth.trajectory(world, 8, step=0.2, seed=4)
    noise = np.random.default_rng(5)
    builder = R.RefBuilder()
    for p in traj[:6]:
        builder.append_scan(p, angles, synth.make_scan(world, p, angles, noise))
    refmap = builder.latest_map()
    nx, ny, mx, my, res = refmap.geometry()
    ones = capi.Grid.from_dense(ctx, np.full((ny, nx), 0.999), mx, my, res, apron=2)
    for _ in range(3):
        capi.Pyramid(ctx, ones, 6).close()
    grid = capi.Grid.from_dense(ctx, refmap.dense(), mx, my, res, apron=2)
    pyr = capi.Pyramid(ctx, grid, 6)
    refpyr = refmap.pyramid(6)
    for lvl in range(7):
        assert np.array_equal(pyr.download(lvl).view(np.int64), refpyr[lvl].dense().view(np.int64))
    bb = dict(node_height_max=6, range_x=2.0, range_y=2.0, range_theta=0.4, scan_range_max=20.0,
              score_range_min=0.01, score_range_max=20.0)
    # the latest map is resized tightly around its scans, so node windows (up to +107 cells) and the
    # beams that end on the outer walls read beyond the map's upper / right edge all the time
    found = 0
    for k, shift in enumerate(([-0.5, -0.4, 0.05], [0.6, 0.3, -0.08], [0.2, -0.7, 0.1])):
        scan = synth.make_scan(world, traj[6], angles, np.random.default_rng(50 + k))
        init = traj[6] + np.array(shift)
        batch = capi.BbBatch(ctx, **bb)
        batch.upload(capi.Scans([angles], [scan], [init], range_min=0.02, range_max=30.0), [pyr], 0.2)
        batch.run()
        (out,) = batch.results()
        ref = R.bb_match(refmap, angles, scan, init, pyramid=refpyr, thr=0.2, height_max=6, range_x=2.0,
                         range_y=2.0, range_theta=0.4, scan_range_max=20.0, score_range_min=0.01,
                         score_range_max=20.0)
        assert (out.found, out.ix, out.iy, out.it) == (ref.found, ref.ix, ref.iy, ref.it)
        if ref.found:
            assert out.score == ref.score
            found += 1
        batch.close()
    assert found >= 2
