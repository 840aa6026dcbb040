"""GPU parity: exhaustive grid-search matcher (ScanMatcherGridSearch + ScorePixelAccurate) through the
C ABI vs the reference's own matcher: winning loop counters, loop lengths and scores bit-identical."""
import os

import numpy as np
import pytest

from my_lidar_graph_slam_b200 import capi, synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

GS_PARAMS = [dict(range_x=0.6, range_y=0.5, range_theta=0.12, step_x=0.05, step_y=0.05, step_theta=0.01),
             dict(range_x=0.4, range_y=0.4, range_theta=0.1, step_x=0.03, step_y=0.07, step_theta=0.013),
             dict(range_x=0.3, range_y=0.3, range_theta=0.05, step_x=0.1, step_y=0.1, step_theta=0.005)]


def _ints(r):
    return [r.found, r.ix, r.iy, r.it, r.win_x, r.win_y, r.win_t]


@pytest.fixture(params=["fused", "tables"], autouse=True)
def gs_path(request, ctx):
    """Every test runs through the fused kernel and through the index-table path ("gs_tables")."""
    ctx.set_option("gs_tables", 1 if request.param == "tables" else 0)
    yield request.param
    ctx.set_option("gs_tables", 0)


def test_gs_golden_vectors(ctx):
    """tests/golden/scene_gs.npz was written by the unmodified reference matcher."""
    g, c = np.load(os.path.join(GOLD, "scene_rooms.npz")), np.load(os.path.join(GOLD, "scene_gs.npz"))
    angles, scans = g["angles"], g["scans"]
    nx, ny, mx, my, res = g["local_geom"][:5]
    grid = capi.Grid.from_dense(ctx, g["local_dense"], mx, my, res, apron=1)
    for pi, p in enumerate(GS_PARAMS):
        sel = [n for n in range(len(c["inits"])) if n % 3 == pi]
        batch = capi.Scans([angles] * len(sel), [scans[c["scan"][n]] for n in sel], c["inits"][sel],
                           range_min=0.02, range_max=30.0)
        out = capi.gs_match(ctx, batch, [grid] * len(sel), norm_threshold=c["thr"][sel], **p)
        for r, n in zip(out, sel):
            assert _ints(r) == list(c["ints"][n])
            if r.found:
                assert r.score == c["flts"][n][0]


def test_gs_every_hypothesis_score_bit_exact(ctx):
    """The whole [theta][y][x] score table against ScorePixelAccurate at the accumulated poses."""
    from oracle import backend
    from scenes import room_scene
    R = backend()
    world, angles, traj, builder = room_scene(seed=1)
    refmap = builder.latest_map()
    nx, ny, mx, my, res = refmap.geometry()
    grid = capi.Grid.from_dense(ctx, refmap.dense(), mx, my, res, apron=1)
    level0 = refmap.pyramid(0)[0]
    scan = synth.make_scan(world, traj[11], angles, np.random.default_rng(9))
    init = traj[11] + np.array([0.07, -0.04, 0.02])
    p = dict(range_x=0.23, range_y=0.31, range_theta=0.04, step_x=0.037, step_y=0.05, step_theta=0.007)
    batch = capi.Scans([angles], [scan], [init], range_min=0.02, range_max=30.0)
    (r,), table = capi.gs_match(ctx, batch, [grid], norm_threshold=0.1, want_table=True, **p)
    dx, dy, dt = (capi.gs_offsets(p["range_x"], p["step_x"]), capi.gs_offsets(p["range_y"], p["step_y"]),
                  capi.gs_offsets(p["range_theta"], p["step_theta"]))
    assert table.shape == (len(dt), len(dy), len(dx)) == (r.win_t, r.win_y, r.win_x)
    kw = dict(geom=refmap) if R.__name__.endswith("portapi") else {}
    want = np.array([[[R.pixel_accurate_score(level0, [init[0] + x, init[1] + y, init[2] + t], angles, scan, **kw)
                       for x in dx] for y in dy] for t in dt])
    assert np.array_equal(table.view(np.int64), want.view(np.int64))
    ref = R.gs_match(refmap, angles, scan, init, thr=0.1, **p)
    assert _ints(r) == [ref.found, ref.ix, ref.iy, ref.it, ref.winX, ref.winY, ref.winT] and r.found
    assert r.score == ref.score == want.max()
    # forced host fix-ups (huge guard band: every beam takes the glibc path) change nothing
    try:
        ctx.set_edge_eps(0.3)
        (r2,), table2 = capi.gs_match(ctx, batch, [grid], norm_threshold=0.1, want_table=True, **p)
    finally:
        ctx.set_edge_eps(0.0)
    assert r2.n_fixups > 1000 and np.array_equal(table2, table) and _ints(r2) == _ints(r)


def test_gs_batch_of_maps_ragged_scans_and_thresholds(ctx):
    from oracle import backend
    R = backend()
    rng = np.random.default_rng(51)
    maps, grids = [], []
    for m in range(3):
        dense = np.where(rng.random((128, 192)) < 0.4, np.round(rng.uniform(1e-3, 0.999, (128, 192)), 3), 0.0)
        maps.append(R.RefMap.from_dense(dense, -3.0 + m, -2.0))
        grids.append(capi.Grid.from_dense(ctx, dense, -3.0 + m, -2.0, 0.05, apron=2))
    beams = [3, 45, 181, 400, 77, 1]
    angles = [np.linspace(-2.0, 2.0, n) for n in beams]
    ranges = [rng.uniform(0.0, 5.0, n) for n in beams]
    inits = np.stack([rng.uniform(-3.5, 6.0, 6), rng.uniform(-2.5, 4.5, 6), rng.uniform(-3, 3, 6)], axis=1)
    which = [0, 1, 2, 0, 1, 2]
    thr = np.array([0.05, 0.1, 0.9, 0.02, 0.3, 1e-300])
    p = dict(range_x=0.5, range_y=0.35, range_theta=0.2, step_x=0.11, step_y=0.05, step_theta=0.03,
             score_range_min=0.3, score_range_max=4.0)
    batch = capi.Scans(angles, ranges, inits, range_min=0.1, range_max=4.5)
    out = capi.gs_match(ctx, batch, [grids[w] for w in which], norm_threshold=thr, **p)
    found = 0
    for q, r in enumerate(out):
        ref = R.gs_match(maps[which[q]], angles[q], ranges[q], inits[q], thr=float(thr[q]), scan_min_range=0.1,
                         scan_max_range=4.5, **p)
        assert _ints(r) == [ref.found, ref.ix, ref.iy, ref.it, ref.winX, ref.winY, ref.winT], q
        if r.found:
            assert r.score == ref.score
            found += 1
    assert 2 <= found < len(out)


def test_gs_edge_cases(ctx):
    dense = np.zeros((64, 64))
    dense[20:40, 20:40] = 0.8
    grid = capi.Grid.from_dense(ctx, dense, 0.0, 0.0, 0.05, apron=1)
    assert capi.gs_offsets(0.0, 0.05).tolist() == [-0.0] and len(capi.gs_offsets(0.1, 1.0)) == 1
    assert len(capi.gs_offsets(2.0, 0.05)) in (40, 41)
    one = capi.Scans([np.array([0.0, 0.1])], [np.array([0.5, 0.5])], [[1.0, 1.5, 0.0]])
    (r,) = capi.gs_match(ctx, one, [grid], range_x=0.0, range_y=0.0, range_theta=0.0, norm_threshold=0.1)
    assert (r.found, r.ix, r.iy, r.it, r.win_x, r.win_y, r.win_t) == (1, 0, 0, 0, 1, 1, 1) and r.score == 1.6
    empty = capi.Scans([np.zeros(0)], [np.zeros(0)], [[1.0, 1.5, 0.0]])
    (r,) = capi.gs_match(ctx, empty, [grid], range_x=0.1, range_y=0.1, range_theta=0.01)
    assert r.found == 0 and (r.ix, r.iy, r.it) == (-1, -1, -1)
    assert capi.gs_match(ctx, capi.Scans([], [], np.zeros((0, 3))), []) == []
    with pytest.raises(capi.LgsError):
        capi.gs_match(ctx, one, [grid], step_x=0.0)


def test_gs_wide_window_uses_the_1024_thread_block(ctx):
    """61 x 51 offsets per theta: 793 threads per block in the fused kernel."""
    from oracle import backend
    R = backend()
    rng = np.random.default_rng(61)
    dense = np.where(rng.random((128, 192)) < 0.45, np.round(rng.uniform(1e-3, 0.999, (128, 192)), 3), 0.0)
    refmap = R.RefMap.from_dense(dense, -3.0, -2.0)
    grid = capi.Grid.from_dense(ctx, dense, -3.0, -2.0, 0.05, apron=1)
    angles = [np.linspace(-1.5, 1.5, 150), np.linspace(-3.0, 3.0, 97)]
    ranges = [rng.uniform(0.2, 3.5, 150), rng.uniform(0.2, 3.5, 97)]
    inits = np.array([[0.5, 0.3, 0.2], [2.0, 1.0, -1.0]])
    p = dict(range_x=1.2, range_y=0.5, range_theta=0.2, step_x=0.02, step_y=0.01, step_theta=0.03)
    batch = capi.Scans(angles, ranges, inits, range_min=0.02, range_max=30.0)
    out = capi.gs_match(ctx, batch, [grid, grid], norm_threshold=0.05, **p)
    for q, r in enumerate(out):
        ref = R.gs_match(refmap, angles[q], ranges[q], inits[q], thr=0.05, **p)
        assert (r.win_x, r.win_y) == (61, 51) or (r.win_x, r.win_y) == (ref.winX, ref.winY)
        assert _ints(r) == [ref.found, ref.ix, ref.iy, ref.it, ref.winX, ref.winY, ref.winT] and r.found
        assert r.score == ref.score
