"""GPU parity: occupancy-grid scan integration vs the reference's GridMapBuilder (every cell bit-exact)."""
import ctypes as C

import numpy as np
import pytest

from my_lidar_graph_slam_b200 import capi, synth

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a).view(np.int64)


def _scene(seed=3, n=40, beams=1081, fov=270.0):
    world = synth.RoomsWorld(40.0, 5.0, seed=seed)
    angles = synth.beam_angles(beams, fov)
    traj = synth.trajectory(world, n, step=0.25, seed=seed)
    noise = np.random.default_rng(seed + 1)
    scans = [synth.make_scan(world, p, angles, noise) for p in traj]
    return angles, traj, scans


def test_incremental_local_map_matches_builder(ctx):
    """UpdateGridMap path: grow (Expand, 5 m slack) + integrate one scan at a time, 40 scans."""
    from oracle import backend
    R = backend()
    angles, traj, scans = _scene()
    builder = R.RefBuilder()
    geo = capi.Geometry(0, 0, traj[0][0], traj[0][1], 0.05, 64)
    grid = capi.Grid(ctx, 0, 0, geo.min_x, geo.min_y, 0.05, apron=2)
    updates = 0
    for k, (p, r) in enumerate(zip(traj, scans)):
        builder.append_scan(p, angles, r)
        hits, bbox = capi.scan_hit_points(p, angles, r, 0.02, 20.0)   # zero relative sensor pose
        geo, sx, sy, changed = capi.geometry_expand(geo, bbox)
        if changed:
            capi.grid_resize(grid, geo, sx, sy)
        updates += capi.integrate_scans(ctx, grid, [p], [hits])
        if k in (0, 7, 39):
            ref = builder.local_map(0)
            assert geo.as_tuple() == ref.geometry()
            assert np.array_equal(_bits(grid.download()), _bits(ref.dense())), f"after scan {k}"
    assert updates > 40 * 50_000


def test_batched_construct_map_matches_latest_map(ctx):
    """ConstructMapFromScans path: tight Resize + Reset + all scans of the window in ONE batch."""
    from oracle import backend
    R = backend()
    angles, traj, scans = _scene(seed=5, n=14)
    builder = R.RefBuilder(n_latest=10)
    for p, r in zip(traj, scans):
        builder.append_scan(p, angles, r)
    ref = builder.latest_map()                     # scans 4..13
    # replay the latest map's geometry history like UpdateLatestMap does every frame
    geo = capi.Geometry(0, 0, 0.0, 0.0, 0.05, 64)
    for last in range(len(traj)):
        lo = max(0, last - 9)
        hp = [capi.scan_hit_points(traj[k], angles, scans[k], 0.02, 20.0) for k in range(lo, last + 1)]
        bl = np.min([[b[0], b[1]] for _, b in hp], axis=0)
        # grid_map_builder.cpp:236-237 seeds the top-right bound with numeric_limits<double>::min()
        tr = np.maximum(np.max([[b[2], b[3]] for _, b in hp], axis=0), np.finfo(np.float64).tiny)
        geo, _, _ = capi.geometry_resize(geo, (bl[0], bl[1], tr[0], tr[1]))
    assert geo.as_tuple() == ref.geometry()
    grid = capi.Grid(ctx, geo.nx, geo.ny, geo.min_x, geo.min_y, 0.05, apron=2)
    capi.integrate_scans(ctx, grid, traj[4:14], [h for h, _ in hp])
    assert np.array_equal(_bits(grid.download()), _bits(ref.dense()))
    # a second pass over the same scans (e.g. AfterLoopClosure re-integration into a Reset map)
    capi.grid_clear(grid)
    capi.integrate_scans(ctx, grid, traj[4:9], [h for h, _ in hp[:5]])
    capi.integrate_scans(ctx, grid, traj[9:14], [h for h, _ in hp[5:]])
    assert np.array_equal(_bits(grid.download()), _bits(ref.dense()))


@pytest.mark.parametrize("case", ["unsorted", "full_circle", "short_rays", "dense_hits"])
def test_adversarial_scans_match_reference_loop(ctx, case):
    """Synthetic hit sets that stress the candidate search: non-monotone beams (fallback path),
    a 360-degree scan (angle wrap), rays shorter than a cell, and many hits in few cells (long
    mixed miss/hit sequences in the near field -> RLE overflow fallback)."""
    from oracle import backend
    R = backend()
    rng = np.random.default_rng({"unsorted": 1, "full_circle": 2, "short_rays": 3, "dense_hits": 4}[case])
    ref = R.RefMap.from_dense(np.zeros((256, 256)), -6.4, -6.4)
    grid = capi.Grid(ctx, 256, 256, -6.4, -6.4, 0.05, apron=1)
    for s in range(6):
        sensor = rng.uniform(-1.5, 1.5, 2)
        n = 700
        if case == "unsorted":
            ang = rng.uniform(-np.pi, np.pi, n)
            rad = rng.uniform(0.05, 4.0, n)
        elif case == "full_circle":
            ang = np.linspace(-np.pi, np.pi, n, endpoint=False) + rng.uniform(-3, 3)
            rad = rng.uniform(0.5, 4.5, n)
        elif case == "short_rays":
            ang = np.linspace(-2.0, 2.0, n)
            rad = np.where(rng.random(n) < 0.5, rng.uniform(0.001, 0.12, n), rng.uniform(0.2, 3.0, n))
        else:
            ang = np.linspace(-1.0, 1.0, n)
            rad = np.where(np.arange(n) % 3 == 0, rng.uniform(0.3, 0.5, n), rng.uniform(0.9, 1.1, n))
        hits = np.ascontiguousarray(sensor[None, :] + rad[:, None] * np.stack([np.cos(ang), np.sin(ang)], 1))
        want = R.map_integrate_hits(ref, sensor, hits)
        got = capi.integrate_scans(ctx, grid, [sensor], [hits])
        assert got == want
        assert np.array_equal(_bits(grid.download()), _bits(ref.dense())), f"{case}: scan {s}"


def test_integrate_rejects_cells_outside_grid_and_handles_empty(ctx):
    grid = capi.Grid(ctx, 64, 64, 0.0, 0.0, 0.05, apron=1)
    assert capi.integrate_scans(ctx, grid, np.zeros((0, 2)), []) == 0
    assert capi.integrate_scans(ctx, grid, [[1.0, 1.0]], [np.zeros((0, 2))]) == 0
    with pytest.raises(capi.LgsError, match="outside"):
        capi.integrate_scans(ctx, grid, [[1.0, 1.0]], [np.array([[5.0, 1.0]])])
    assert not grid.download().any()


def test_real_scans_never_need_the_exhaustive_fallback(ctx):
    """Touch sequences that fit no 32-bit record go to the side buffer; only when that is exhausted
    does the owning thread re-derive a (cell, scan) from the beams.  On real scan geometry the
    default side buffer must always suffice."""
    angles, traj, scans = _scene(seed=9, n=24)
    geo = capi.Geometry(0, 0, traj[0][0], traj[0][1], 0.05, 64)
    hits = []
    for p, r in zip(traj, scans):
        h, bbox = capi.scan_hit_points(p, angles, r, 0.02, 20.0)
        geo, _, _, _ = capi.geometry_expand(geo, bbox)
        hits.append(h)
    grid = capi.Grid(ctx, geo.nx, geo.ny, geo.min_x, geo.min_y, 0.05, apron=1)
    before = capi.lib().lgs_ctx_integrate_fallback_cells(ctx.h)
    assert capi.integrate_scans(ctx, grid, traj[:, :2], hits) > 24 * 50_000
    assert capi.lib().lgs_ctx_integrate_fallback_cells(ctx.h) == before


def test_exhausted_side_buffer_falls_back_to_exact_beam_tests(ctx):
    """The "integ_side_words" option shrinks the side buffer to nothing: every long mixed sequence takes the
    exhaustive per-beam path in the fold pass and the map must still be bit-identical."""
    from oracle import backend
    R = backend()
    angles, traj, scans = _scene(seed=11, n=12)
    geo = capi.Geometry(0, 0, traj[0][0], traj[0][1], 0.05, 64)
    hits = []
    for p, r in zip(traj, scans):
        h, bbox = capi.scan_hit_points(p, angles, r, 0.02, 20.0)
        geo, _, _, _ = capi.geometry_expand(geo, bbox)
        hits.append(h)
    ref = R.RefMap.from_dense(np.zeros((geo.ny, geo.nx)), geo.min_x, geo.min_y)
    want = sum(R.map_integrate_hits(ref, p[:2], h) for p, h in zip(traj, hits))
    maps = []
    for side in ("8", None):
        ctx.set_option("integ_side_words", int(side) if side else 0)
        grid = capi.Grid(ctx, geo.nx, geo.ny, geo.min_x, geo.min_y, 0.05, apron=1)
        before = capi.lib().lgs_ctx_integrate_fallback_cells(ctx.h)
        assert capi.integrate_scans(ctx, grid, traj[:, :2], hits) == want
        used_fallback = capi.lib().lgs_ctx_integrate_fallback_cells(ctx.h) - before
        maps.append((grid.download(), used_fallback))
        grid.close()
    assert maps[0][1] > 0 and maps[1][1] == 0
    assert np.array_equal(_bits(maps[0][0]), _bits(ref.dense()))
    assert np.array_equal(_bits(maps[1][0]), _bits(ref.dense()))


def test_many_chunks_overlapped_fold_matches_reference(ctx):
    """150 scans in ONE call = three 64-scan chunks whose fold passes run on a second stream under the
    next chunk's mark/touch passes (double-buffered records): every cell must still be bit-identical,
    and so must a second call that re-integrates on top of the first (no stale workspace state)."""
    from oracle import backend
    R = backend()
    world = synth.RoomsWorld(40.0, 5.0, seed=21)
    angles = synth.beam_angles(721, 240.0)
    traj = synth.trajectory(world, 150, step=0.12, seed=21)
    noise = np.random.default_rng(22)
    geo = capi.Geometry(0, 0, traj[0][0], traj[0][1], 0.05, 64)
    hits = []
    for p in traj:
        h, bbox = capi.scan_hit_points(p, angles, synth.make_scan(world, p, angles, noise), 0.02, 20.0)
        geo, _, _, _ = capi.geometry_expand(geo, bbox)
        hits.append(h)
    ref = R.RefMap.from_dense(np.zeros((geo.ny, geo.nx)), geo.min_x, geo.min_y)
    grid = capi.Grid(ctx, geo.nx, geo.ny, geo.min_x, geo.min_y, 0.05, apron=1)
    for rep in range(2):
        want = sum(R.map_integrate_hits(ref, p[:2], h) for p, h in zip(traj, hits))
        assert capi.integrate_scans(ctx, grid, traj[:, :2], hits) == want
        assert np.array_equal(_bits(grid.download()), _bits(ref.dense())), f"pass {rep}"
    # a call with empty scans in the middle and a one-beam scan
    mixed = [hits[0], np.zeros((0, 2)), hits[1][:1], hits[2]]
    pos = [traj[0, :2], traj[1, :2], traj[1, :2], traj[2, :2]]
    want = sum(R.map_integrate_hits(ref, p, h) for p, h in zip(pos, mixed) if len(h))
    assert capi.integrate_scans(ctx, grid, np.asarray(pos), mixed) == want
    assert np.array_equal(_bits(grid.download()), _bits(ref.dense()))


def test_very_long_rays_use_the_integer_division_path_and_split_chunks(ctx):
    """Rays longer than 2047 cells leave the float-reciprocal fast path of the Bresenham closed form,
    and their huge tile bounding boxes make the record budget split the call into small chunks."""
    from oracle import backend
    R = backend()
    rng = np.random.default_rng(31)
    nx, ny = 4480, 448                                   # 224 m x 22.4 m at 0.05 m
    ref = R.RefMap.from_dense(np.zeros((ny, nx)), 0.0, 0.0)
    grid = capi.Grid(ctx, nx, ny, 0.0, 0.0, 0.05, apron=1)
    sensors, hits = [], []
    for s in range(9):
        sensor = np.array([rng.uniform(1.0, 12.0), rng.uniform(8.0, 14.0)])
        n = 160
        ang = np.sort(rng.uniform(-0.04, 0.04, n))
        rad = np.where(rng.random(n) < 0.7, rng.uniform(110.0, 205.0, n), rng.uniform(0.3, 40.0, n))
        h = sensor[None, :] + rad[:, None] * np.stack([np.cos(ang), np.sin(ang)], 1)
        h[:, 0] = np.clip(h[:, 0], 0.01, nx * 0.05 - 0.01)
        h[:, 1] = np.clip(h[:, 1], 0.01, ny * 0.05 - 0.01)
        sensors.append(sensor)
        hits.append(np.ascontiguousarray(h))
    want = sum(R.map_integrate_hits(ref, p, h) for p, h in zip(sensors, hits))
    assert capi.integrate_scans(ctx, grid, np.asarray(sensors), hits) == want
    assert want > 9 * 160 * 1500
    assert np.array_equal(_bits(grid.download()), _bits(ref.dense()))


def test_grid_copy_between_contexts_takes_geometry_and_cells(ctx):
    """lgs_grid_copy: the hand-over of a device-resident map (different context, different apron)."""
    rng = np.random.default_rng(77)
    dense = np.where(rng.random((70, 90)) < 0.4, rng.uniform(1e-3, 0.999, (70, 90)), 0.0)
    src = capi.Grid.from_dense(ctx, dense, -1.5, 2.25, 0.05, apron=1)
    other = capi.Context(0)
    dst = capi.Grid(other, 10, 20, 0.0, 0.0, 0.05, apron=17)          # other size, placement and apron
    capi.grid_copy(src, dst)
    assert np.array_equal(dst.download(), dense)
    info = [C.c_int(), C.c_int(), C.c_double(), C.c_double(), C.c_double(), C.c_int()]
    other.check(capi.lib().lgs_grid_info(dst.h, *[C.byref(v) for v in info]))
    assert [v.value for v in info] == [90, 70, -1.5, 2.25, 0.05, 17]
    # the apron stays zero: a 5-cell window max over the copy equals the one over the source
    assert np.array_equal(dst.precompute(5).download(), src.precompute(5).download())
    dense2 = np.round(dense, 1)
    src.upload(dense2)
    capi.grid_copy(src, dst)                                           # same size: plain copy
    assert np.array_equal(dst.download(), dense2)
    assert np.array_equal(capi.grid_download_region(dst, 7, 11, 40, 23), dense2[11:34, 7:47])
    assert capi.grid_download_region(dst, 89, 69, 1, 1)[0, 0] == dense2[69, 89]
    with pytest.raises(capi.LgsError):
        capi.grid_download_region(dst, 80, 0, 20, 5)                   # runs off the right edge
    bad = capi.Grid(other, 90, 70, -1.5, 2.25, 0.1, apron=1)
    with pytest.raises(capi.LgsError):
        capi.grid_copy(src, bad)                                       # resolution mismatch
    other.close()


def test_streamed_submit_wait_matches_reference(ctx):
    """lgs_grid_integrate_submit / _wait: calls of 40 scans with two in flight (the staging of call k + 1
    runs on the copy stream under the passes of call k, records double buffered across calls): update
    counts per call, every cell, and the rules of the pair (third submit refused, a bad call leaves the
    grid alone, the synchronous call drains what is in flight)."""
    from oracle import backend
    R = backend()
    world = synth.RoomsWorld(40.0, 5.0, seed=31)
    angles = synth.beam_angles(541, 240.0)
    traj = synth.trajectory(world, 200, step=0.12, seed=31)
    noise = np.random.default_rng(32)
    geo = capi.Geometry(0, 0, traj[0][0], traj[0][1], 0.05, 64)
    hits = []
    for p in traj:
        h, bbox = capi.scan_hit_points(p, angles, synth.make_scan(world, p, angles, noise), 0.02, 20.0)
        geo, _, _, _ = capi.geometry_expand(geo, bbox)
        hits.append(h)
    ref = R.RefMap.from_dense(np.zeros((geo.ny, geo.nx)), geo.min_x, geo.min_y)
    grid = capi.Grid(ctx, geo.nx, geo.ny, geo.min_x, geo.min_y, 0.05, apron=1)
    calls = [capi.PackedHits(traj[k:k + 40, :2], hits[k:k + 40]) for k in range(0, 200, 40)]
    want = [sum(R.map_integrate_hits(ref, p[:2], h) for p, h in zip(traj[k:k + 40], hits[k:k + 40]))
            for k in range(0, 200, 40)]
    got, in_flight = [], 0
    for c in calls:
        capi.integrate_submit(ctx, grid, c)
        in_flight += 1
        if in_flight == 2:
            if len(got) == 0:                                        # two in flight: a third is refused
                with pytest.raises(capi.LgsError, match="in flight"):
                    capi.integrate_submit(ctx, grid, calls[0])
            got.append(capi.integrate_wait(ctx))
            in_flight -= 1
    while in_flight:
        got.append(capi.integrate_wait(ctx))
        in_flight -= 1
    assert got == want
    assert capi.integrate_wait(ctx) == 0                             # nothing in flight
    assert np.array_equal(_bits(grid.download()), _bits(ref.dense()))
    # a call that leaves the grid fails at submit and changes nothing; one in flight is drained by the
    # synchronous call that follows
    before = grid.download()
    with pytest.raises(capi.LgsError, match="outside"):
        capi.integrate_submit(ctx, grid, capi.PackedHits([[traj[0][0], traj[0][1]]], [np.array([[1e4, 0.0]])]))
    assert np.array_equal(_bits(grid.download()), _bits(before))
    capi.integrate_submit(ctx, grid, calls[0])
    more = sum(R.map_integrate_hits(ref, p[:2], h) for p, h in zip(traj[:40], hits[:40]))
    more2 = sum(R.map_integrate_hits(ref, p[:2], h) for p, h in zip(traj[40:80], hits[40:80]))
    assert capi.integrate_packed(ctx, grid, calls[1]) == more2 and more > 0
    assert np.array_equal(_bits(grid.download()), _bits(ref.dense()))
