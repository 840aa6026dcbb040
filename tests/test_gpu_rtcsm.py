"""GPU parity: real-time correlative matcher through the C ABI vs the reference's own matcher.

Bar (BASELINE.json north_star): winning (ix, iy, itheta) and projected cell indices bit-exact;
scores within 1e-5 relative (they are in fact bit-identical here, asserted as such)."""
import numpy as np
import pytest

from my_lidar_graph_slam_b200 import capi, synth

pytestmark = pytest.mark.gpu

C2 = dict(low_res=5, range_x=1.0, range_y=1.0, range_theta=1.0471975512, scan_range_max=5.7296)


def _match_inputs(scene, k, perturb_seed=3):
    world, angles, traj, builder = scene
    rng = np.random.default_rng(perturb_seed + k)
    true = traj[10 + (k % 6)]
    scan = synth.make_scan(world, true, angles, np.random.default_rng(100 + k))
    init = true + np.array([rng.uniform(-0.3, 0.3), rng.uniform(-0.3, 0.3), rng.uniform(-0.2, 0.2)])
    return scan, init


def _device_maps(ctx, refmap, win, apron=32):
    nx, ny, mx, my, res = refmap.geometry()
    g = capi.Grid.from_dense(ctx, refmap.dense(), mx, my, res, apron=apron)
    return g, g.precompute(win)


def test_rtcsm_c2_winner_and_tables_bit_exact(ctx):
    from oracle import backend
    R = backend()
    from scenes import room_scene
    scene = room_scene(seed=1)
    world, angles, traj, builder = scene
    refmap = builder.latest_map()
    pre = refmap.precompute(5)
    grid, coarse = _device_maps(ctx, refmap, 5)
    scan, init = _match_inputs(scene, 0)
    ref = R.rtcsm_match(refmap, angles, scan, init, pre=pre, **C2)

    batch = capi.RtcsmBatch(ctx, **C2)
    scans = capi.Scans([angles], [scan], [init])     # zero relative sensor pose
    batch.upload(grid, scans)
    batch.run(grid, coarse)
    (out,) = batch.results(grid, coarse)
    assert (out.win_x, out.win_y, out.win_t) == (ref.winX, ref.winY, ref.winT)
    assert out.step_t == ref.stepT
    assert (out.found, out.ix, out.iy, out.it) == (ref.found, ref.ix, ref.iy, ref.it)
    assert out.score == ref.score                      # bit-identical, not just 1e-5

    # every projected cell index and every fine / coarse score against the reference's own
    # ComputeScanIndices / ComputeScore
    fine, coarse_t, cells = batch.debug(0)
    nt, nyw, nxw = fine.shape
    tab, idx, cnt = R.rtcsm_score_table(refmap, pre, False, 5, C2["scan_range_max"], init, angles,
                                        scan, ref.stepT, ref.winT, -ref.winX, nxw, -ref.winY, nyw)
    assert (cnt == cells.shape[1]).all()
    assert np.array_equal(idx[:, :cells.shape[1], :], cells)
    assert np.array_equal(tab.view(np.int64), fine.view(np.int64))
    nbx, nby = coarse_t.shape[1:]
    ctab, _, _ = R.rtcsm_score_table(refmap, pre, True, 5, C2["scan_range_max"], init, angles, scan,
                                     ref.stepT, ref.winT, -ref.winX, nxw, -ref.winY, nyw)
    ref_coarse = ctab[:, ::5, ::5].transpose(0, 2, 1)   # [t][bx][by]
    assert np.array_equal(ref_coarse.view(np.int64), coarse_t.view(np.int64))


def test_rtcsm_batch_matches_reference(ctx):
    from oracle import backend
    R = backend()
    from scenes import room_scene
    scene = room_scene(seed=2)
    world, angles, traj, builder = scene
    refmap = builder.latest_map()
    pre = refmap.precompute(5)
    grid, coarse = _device_maps(ctx, refmap, 5)
    n = 12
    ins = [_match_inputs(scene, k) for k in range(n)]
    batch = capi.RtcsmBatch(ctx, **C2)
    batch.upload(grid, capi.Scans([angles] * n, [s for s, _ in ins], [p for _, p in ins]))
    batch.run(grid, coarse)
    outs = batch.results(grid, coarse)
    for (scan, init), out in zip(ins, outs):
        ref = R.rtcsm_match(refmap, angles, scan, init, pre=pre, **C2)
        assert (out.found, out.ix, out.iy, out.it) == (ref.found, ref.ix, ref.iy, ref.it)
        assert out.score == ref.score


@pytest.mark.parametrize("params,thr", [
    (dict(low_res=5, range_x=0.2, range_y=0.2, range_theta=0.5, scan_range_max=20.0), None),  # C1 defaults
    (dict(low_res=3, range_x=0.6, range_y=0.4, range_theta=0.3, scan_range_max=8.0), 0.3),
    (dict(low_res=1, range_x=0.3, range_y=0.3, range_theta=0.1, scan_range_max=20.0), 0.95),  # not found
    (dict(low_res=7, range_x=1.5, range_y=0.7, range_theta=0.2, scan_range_max=10.0), 0.5),
])
def test_rtcsm_parameter_sweep(ctx, params, thr):
    from oracle import backend
    R = backend()
    from scenes import room_scene
    scene = room_scene(seed=3, n_beams=361, fov=180.0)
    world, angles, traj, builder = scene
    refmap = builder.latest_map()
    pre = refmap.precompute(params["low_res"])
    grid, coarse = _device_maps(ctx, refmap, params["low_res"], apron=48)
    for k in range(3):
        scan, init = _match_inputs(scene, k)
        ref = R.rtcsm_match(refmap, angles, scan, init, pre=pre, thr=thr, **params)
        batch = capi.RtcsmBatch(ctx, **params)
        batch.upload(grid, capi.Scans([angles], [scan], [init]), None if thr is None else [thr])
        batch.run(grid, coarse)
        (out,) = batch.results(grid, coarse)
        assert (out.win_x, out.win_y, out.win_t) == (ref.winX, ref.winY, ref.winT)
        assert (out.found, out.ix, out.iy, out.it) == (ref.found, ref.ix, ref.iy, ref.it)
        if ref.found:
            assert out.score == ref.score


def test_rtcsm_scan_overhanging_lower_left_edge(ctx):
    """H12: scans hanging over the lower-left map edge, where coarse values are not bounds."""
    from oracle import backend
    R = backend()
    rng = np.random.default_rng(7)
    ny, nx = 128, 128
    dense = np.where(rng.random((ny, nx)) < 0.25, rng.uniform(0.05, 0.95, (ny, nx)), 0.0)
    dense[:12, :] = rng.uniform(0.5, 0.99, (12, nx))     # strong structure right at the low edges
    dense[:, :12] = rng.uniform(0.5, 0.99, (ny, 12))
    refmap = R.RefMap.from_dense(dense, -1.0, -2.0)
    pre = refmap.precompute(5)
    grid, coarse = _device_maps(ctx, refmap, 5)
    assert np.array_equal(grid.download(), refmap.dense())
    params = dict(low_res=5, range_x=1.0, range_y=1.0, range_theta=0.2, scan_range_max=20.0)
    angles = synth.beam_angles(181, 180.0)
    replayed = differs_from_exhaustive = 0
    for k in range(10):
        # short beams pointing at the x < 0 half plane from just inside the left edge: for many
        # window offsets the coarse index is negative (reads 0) while the fine cells are inside
        ranges = rng.uniform(0.1, 0.5, angles.shape)
        init = np.array([-1.0 + rng.uniform(0.0, 0.3), -2.0 + rng.uniform(0.5, 3.0),
                         np.pi + rng.uniform(-0.3, 0.3)])
        ref = R.rtcsm_match(refmap, angles, ranges, init, pre=pre, **params)
        batch = capi.RtcsmBatch(ctx, **params)
        batch.upload(grid, capi.Scans([angles], [ranges], [init]))
        batch.run(grid, coarse)
        out, = batch.results(grid, coarse)
        assert (out.found, out.ix, out.iy, out.it) == (ref.found, ref.ix, ref.iy, ref.it)
        assert out.score == ref.score
        replayed += out.exact_replay
        fine, _, _ = batch.debug(0)
        differs_from_exhaustive += int(fine.max() != ref.score)
    assert replayed > 0, "scene was meant to exercise the sequential CPU-order replay"
    # in this scene the reference's pruned search really does miss the exhaustive optimum
    assert differs_from_exhaustive > 0


def _run_one(ctx, grid, coarse, params, angles, ranges, init, thr=None):
    batch = capi.RtcsmBatch(ctx, **params)
    batch.upload(grid, capi.Scans([angles], [ranges], [init]), thr)
    batch.run(grid, coarse)
    return batch.results(grid, coarse)


def test_rtcsm_empty_and_degenerate(ctx):
    from oracle import backend
    R = backend()
    dense = np.zeros((64, 64))
    dense[10:20, 30] = 0.9
    refmap = R.RefMap.from_dense(dense, 0.0, 0.0)
    grid, coarse = _device_maps(ctx, refmap, 5)
    params = dict(low_res=5, range_x=0.5, range_y=0.5, range_theta=0.1, scan_range_max=4.0)
    angles = synth.beam_angles(91, 90.0)
    # every beam beyond scanRangeMax -> no kept beams -> nothing found, initial indices returned
    ranges = np.full(angles.shape, 6.0)
    init = np.array([1.0, 1.0, 0.3])
    ref = R.rtcsm_match(refmap, angles, ranges, init, **params)
    out, = _run_one(ctx, grid, coarse, params, angles, ranges, init)
    assert ref.found == 0 and out.found == 0
    assert (out.ix, out.iy, out.it) == (ref.ix, ref.iy, ref.it) == (-out.win_x, -out.win_y, -out.win_t)
    # empty batch
    batch = capi.RtcsmBatch(ctx, **params)
    batch.upload(grid, capi.Scans([], [], np.zeros((0, 3))))
    batch.run(grid, coarse)
    assert batch.results(grid, coarse) == []
    # window wider than the apron is refused, not silently clipped
    wide = capi.RtcsmBatch(ctx, low_res=5, range_x=10.0, range_y=10.0, range_theta=0.1, scan_range_max=4.0)
    with pytest.raises(capi.LgsError, match="APRON"):
        wide.upload(grid, capi.Scans([angles], [ranges], [init]))
