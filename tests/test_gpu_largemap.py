"""GPU parity: a large map split into row bands (config C5) gives bit-identical branch-and-bound
results and pyramid values to the whole map on one grid, and to the reference matcher."""
import numpy as np
import pytest

from my_lidar_graph_slam_b200 import capi, largemap, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["device", "exact"])
def bb_run_path(request, ctx):
    """Every test runs twice: through the device-only run (one persistent kernel, fixed-point cells,
    near-edge points decided on the device) and through the level-synchronous exact path (full index
    table, near-edge points from the host)."""
    ctx.set_option("bb_sync", 1 if request.param == "exact" else 0)
    yield request.param
    ctx.set_option("bb_sync", 0)

# short usable range so that three bands of a ~900-row map are real windows (margins ~100 / ~230 rows)
P = dict(node_height_max=6, range_x=2.0, range_y=2.0, range_theta=0.5, scan_range_max=4.0,
         score_range_min=0.01, score_range_max=4.0)


@pytest.fixture(scope="module")
def scene(ctx):
    from oracle import backend
    R = backend()
    world = synth.RoomsWorld(40.0, 5.0, seed=8)
    angles = synth.beam_angles(1081, 270.0)
    # four short walks in rooms spread along y, so that every row band owns some of the queries
    traj = np.concatenate([synth.trajectory(world, 12, step=0.4, seed=8 + k, start=(-2.5, y, 0.3 * k))
                           for k, y in enumerate((-12.5, -2.5, 2.5, 12.5))])
    noise = np.random.default_rng(3)
    builder = R.RefBuilder()
    for p in traj:
        builder.append_scan(p, angles, synth.make_scan(world, p, angles, noise))
    refmap = builder.local_map(0)
    return dict(world=world, angles=angles, traj=traj, refmap=refmap, R=R)


def _queries(sc, n):
    rng = np.random.default_rng(17)
    qs = []
    for k in range(n):
        true = sc["traj"][int(rng.integers(0, len(sc["traj"])))]
        scan = synth.make_scan(sc["world"], true, sc["angles"], np.random.default_rng(500 + k))
        # beams longer than score_range_max = 4 m are skipped by the score function, so 4 m is the reach
        init = true + np.array([rng.uniform(-0.4, 0.4), rng.uniform(-0.4, 0.4), rng.uniform(-0.1, 0.1)])
        qs.append((scan, init))
    return qs


def _match(ctx, pyramids, angles, qs):
    batch = capi.BbBatch(ctx, **P)
    scans = capi.Scans([angles] * len(qs), [s for s, _ in qs], [p for _, p in qs], range_min=0.02, range_max=30.0)
    batch.upload(scans, pyramids, 0.4)
    batch.run()
    out = batch.results()
    batch.close()
    return out


@pytest.mark.parametrize("world", [2, 3])
def test_banded_map_matches_whole_map_and_reference(ctx, scene, world):
    nx, ny, mx, my, res = scene["refmap"].geometry()
    dense = scene["refmap"].dense()
    whole = capi.Grid.from_dense(ctx, dense, mx, my, res, apron=1)
    whole_pyr = capi.Pyramid(ctx, whole, 6)
    qs = _queries(scene, 18)
    want = _match(ctx, [whole_pyr] * len(qs), scene["angles"], qs)

    bands = [largemap.BandedMap(ctx, lambda a, b: dense[a:b], nx, ny, mx, my, res, g, world, 6,
                                reach_m=4.0, range_y_m=P["range_y"]) for g in range(world)]
    assert any(b.w0 > 0 for b in bands) and any(b.w1 < ny for b in bands)     # real windows
    for b in bands:
        b.build_pyramid()
        # pyramid rows that a match may read equal the whole map's (the top 2^H - 1 window rows of a
        # band that ends inside the map are clamped differently and are never read)
        safe = (b.w1 - b.w0) if b.w1 == ny else (b.w1 - b.w0) - 63
        for lvl in (1, 3, 6):
            assert np.array_equal(b.pyramid.download(lvl)[:safe].view(np.int64),
                                  whole_pyr.download(lvl)[b.w0:b.w0 + safe].view(np.int64))
    rows = largemap.sensor_rows([p[1] for _, p in qs], my, res)
    owner = largemap.owner_of_rows(rows, ny, world)
    assert len(set(owner.tolist())) == world                                   # every band gets work
    got = [None] * len(qs)
    for g, b in enumerate(bands):
        mine = np.flatnonzero(owner == g)
        outs = _match(ctx, [b.pyramid] * len(mine), scene["angles"], [qs[k] for k in mine])
        for k, o in zip(mine, outs):
            got[k] = o
    found = 0
    for k, (a, b_) in enumerate(zip(got, want)):
        assert (a.found, a.ix, a.iy, a.it) == (b_.found, b_.ix, b_.iy, b_.it), k
        assert a.score == b_.score
        found += a.found
    assert found >= 4
    # and the reference's own matcher on the whole map agrees (first few queries: it is slow)
    R = scene["R"]
    refpyr = scene["refmap"].pyramid(6)
    for k in range(4):
        ref = R.bb_match(scene["refmap"], scene["angles"], qs[k][0], qs[k][1], pyramid=refpyr, thr=0.4,
                         height_max=6, range_x=P["range_x"], range_y=P["range_y"], range_theta=P["range_theta"],
                         scan_range_max=P["scan_range_max"], score_range_min=P["score_range_min"],
                         score_range_max=P["score_range_max"])
        assert (got[k].found, got[k].ix, got[k].iy, got[k].it) == (ref.found, ref.ix, ref.iy, ref.it)
        if ref.found:
            assert got[k].score == ref.score
    for b in bands:
        b.close()


def test_windowed_grids_are_rejected_where_unsupported(ctx):
    g = capi.Grid(ctx, 64, 64, 0.0, 0.0, 0.05, apron=1)
    g.set_window(0, 16)
    with pytest.raises(capi.LgsError, match="window"):
        capi.integrate_scans(ctx, g, [[1.0, 1.0]], [np.array([[1.5, 1.0]])])
