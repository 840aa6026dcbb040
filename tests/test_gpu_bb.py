"""GPU parity: branch-and-bound matcher through the C ABI vs the reference's ScanMatcherBranchBound.

Winning (ix, iy, itheta) and found flag bit-exact; score bit-identical (bar: 1e-5 relative)."""
import numpy as np
import pytest

from my_lidar_graph_slam_b200 import capi, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["table", "slots"])
def bb_root_path(request, monkeypatch):
    """Every test runs twice: with the library's own choice between the full per-query index table and
    the root-from-hit-points path (queries here mostly have their own scan -> table), and with the
    hit-point path forced."""
    if request.param == "slots":
        monkeypatch.setenv("LGS_BB_SLOTS", "1")
    else:
        monkeypatch.delenv("LGS_BB_SLOTS", raising=False)
    yield request.param

DEF = dict(node_height_max=6, range_x=2.0, range_y=2.0, range_theta=1.0, scan_range_max=20.0,
           score_range_min=0.01, score_range_max=20.0)


def _ref_kwargs(p):
    return dict(height_max=p["node_height_max"], range_x=p["range_x"], range_y=p["range_y"],
                range_theta=p["range_theta"], scan_range_max=p["scan_range_max"],
                score_range_min=p["score_range_min"], score_range_max=p["score_range_max"])


@pytest.fixture(scope="module")
def submap(ctx):
    """A reference-built local map (40 scans) + reference and device pyramids."""
    from oracle import backend
    R = backend()
    world = synth.RoomsWorld(40.0, 5.0, seed=3)
    angles = synth.beam_angles(1081, 270.0)
    traj = synth.trajectory(world, 60, step=0.25, seed=3)
    noise = np.random.default_rng(2)
    builder = R.RefBuilder()
    for p in traj[:40]:
        builder.append_scan(p, angles, synth.make_scan(world, p, angles, noise))
    refmap = builder.local_map(0)
    nx, ny, mx, my, res = refmap.geometry()
    grid = capi.Grid.from_dense(ctx, refmap.dense(), mx, my, res, apron=2)
    return dict(world=world, angles=angles, traj=traj, noise=noise, refmap=refmap,
                refpyr=refmap.pyramid(6), grid=grid, pyr=capi.Pyramid(ctx, grid, 6))


def _queries(sm, n, seed=5):
    rng = np.random.default_rng(seed)
    out = []
    for k in range(n):
        true = sm["traj"][int(rng.integers(5, 58))]
        scan = synth.make_scan(sm["world"], true, sm["angles"], np.random.default_rng(200 + k))
        init = true + np.array([rng.uniform(-0.6, 0.6), rng.uniform(-0.6, 0.6), rng.uniform(-0.3, 0.3)])
        out.append((scan, init))
    return out


def _same(out, ref):
    assert (out.found, out.ix, out.iy, out.it) == (ref.found, ref.ix, ref.iy, ref.it)
    if ref.found:
        assert out.score == ref.score


def test_bb_batch_matches_reference(ctx, submap):
    from oracle import backend
    R = backend()
    qs = _queries(submap, 8)
    batch = capi.BbBatch(ctx, **DEF)
    scans = capi.Scans([submap["angles"]] * len(qs), [s for s, _ in qs], [p for _, p in qs],
                       range_min=0.02, range_max=30.0)
    batch.upload(scans, [submap["pyr"]] * len(qs), 0.6)
    batch.run()
    outs = batch.results()
    found = 0
    for (scan, init), out in zip(qs, outs):
        ref = R.bb_match(submap["refmap"], submap["angles"], scan, init, pyramid=submap["refpyr"],
                         thr=0.6, **_ref_kwargs(DEF))
        assert (out.win_x, out.win_y, out.win_t) == (ref.winX, ref.winY, ref.winT)
        assert out.step_t == ref.stepT
        _same(out, ref)
        found += ref.found
    assert found >= 4
    levels, gathers = batch.work()
    assert levels[6] == sum(2 * o.win_t + 1 for o in outs) and gathers > 0
    # the sequential CPU-order replay must agree with the verified fast path
    batch.force_replay(True)
    batch.run()
    for a, b in zip(outs, batch.results()):
        assert (a.found, a.ix, a.iy, a.it, a.score) == (b.found, b.ix, b.iy, b.it, b.score)
        assert b.exact_replay == 1


@pytest.mark.parametrize("params,thr", [
    (dict(DEF, node_height_max=4, range_x=1.0, range_y=1.5, range_theta=0.4), 0.5),
    (dict(DEF, node_height_max=3, range_x=0.7, range_y=0.7, range_theta=0.2, scan_range_max=8.0,
          score_range_max=8.0), 0.3),
    (dict(DEF, node_height_max=0, range_x=0.3, range_y=0.3, range_theta=0.05), 0.2),
    (dict(DEF, node_height_max=5, range_x=1.0, range_y=1.0, range_theta=0.3), 0.97),   # nothing found
    (dict(DEF, node_height_max=2, range_x=0.4, range_y=0.4, range_theta=0.05), None),  # DBL_MIN: full tree
])
def test_bb_parameter_sweep(ctx, submap, params, thr):
    from oracle import backend
    R = backend()
    H = params["node_height_max"]
    pyr = capi.Pyramid(ctx, submap["grid"], H)
    refpyr = submap["refmap"].pyramid(H)
    for scan, init in _queries(submap, 3, seed=9):
        ref = R.bb_match(submap["refmap"], submap["angles"], scan, init, pyramid=refpyr,
                         thr=thr if thr is not None else float(np.finfo(np.float64).tiny),
                         **_ref_kwargs(params))
        batch = capi.BbBatch(ctx, **params)
        batch.upload(capi.Scans([submap["angles"]], [scan], [init], range_min=0.02, range_max=30.0),
                     [pyr], None if thr is None else thr)
        batch.run()
        (out,) = batch.results()
        _same(out, ref)


def test_bb_near_edge_fixups_do_not_change_results(ctx, submap):
    """Widen the edge guard band so ~5 % of the projected points take the host-exact path."""
    from oracle import backend
    R = backend()
    qs = _queries(submap, 2, seed=21)
    capi.set_edge_eps(0.025)
    try:
        for scan, init in qs:
            ref = R.bb_match(submap["refmap"], submap["angles"], scan, init, pyramid=submap["refpyr"],
                             thr=0.6, **_ref_kwargs(DEF))
            batch = capi.BbBatch(ctx, **DEF)
            batch.upload(capi.Scans([submap["angles"]], [scan], [init], range_min=0.02, range_max=30.0),
                         [submap["pyr"]], 0.6)
            batch.run()
            (out,) = batch.results()
            assert out.n_fixups > 1000
            _same(out, ref)
    finally:
        capi.set_edge_eps(1e-9)


def test_bb_low_edge_overhang_uses_replay(ctx):
    """H12: window indices straddling zero make the win-max values non-bounds; the result must
    still be the reference's (order-dependent) answer."""
    from oracle import backend
    R = backend()
    rng = np.random.default_rng(17)
    ny, nx = 128, 192
    dense = np.where(rng.random((ny, nx)) < 0.25, rng.uniform(0.05, 0.95, (ny, nx)), 0.0)
    dense[:16, :] = rng.uniform(0.5, 0.99, (16, nx))
    dense[:, :16] = rng.uniform(0.5, 0.99, (ny, 16))
    refmap = R.RefMap.from_dense(dense, -1.0, -2.0)
    params = dict(DEF, node_height_max=4, range_x=1.0, range_y=1.0, range_theta=0.2)
    refpyr = refmap.pyramid(4)
    grid = capi.Grid.from_dense(ctx, refmap.dense(), -1.0, -2.0, 0.05, apron=1)
    pyr = capi.Pyramid(ctx, grid, 4)
    angles = synth.beam_angles(181, 180.0)
    replays = 0
    for k in range(12):
        ranges = rng.uniform(0.1, 0.6, angles.shape)
        init = np.array([-1.0 + rng.uniform(0.0, 0.4), -2.0 + rng.uniform(0.3, 3.0),
                         np.pi + rng.uniform(-0.3, 0.3)])
        thr = float(rng.uniform(0.2, 0.5))
        ref = R.bb_match(refmap, angles, ranges, init, pyramid=refpyr, thr=thr, **_ref_kwargs(params))
        batch = capi.BbBatch(ctx, **params)
        batch.upload(capi.Scans([angles], [ranges], [init], range_min=0.02, range_max=30.0), [pyr], thr)
        batch.run()
        (out,) = batch.results()
        _same(out, ref)
        replays += out.exact_replay
    assert replays > 0


def test_bb_empty_batch_and_bad_pyramid(ctx, submap):
    batch = capi.BbBatch(ctx, **DEF)
    batch.upload(capi.Scans([], [], np.zeros((0, 3))), [], None)
    batch.run()
    assert batch.results() == []
    shallow = capi.Pyramid(ctx, submap["grid"], 2)
    scan, init = _queries(submap, 1)[0]
    with pytest.raises(capi.LgsError, match="INVALID"):
        batch.upload(capi.Scans([submap["angles"]], [scan], [init]), [shallow], 0.6)


def test_speculative_sync_free_runs_match_the_level_synchronous_run(ctx, submap, monkeypatch):
    """The first run of a batch object is level-synchronous; later runs launch every level without
    waiting for its node count (pool capacities + device-side counts) and are validated afterwards.
    Both must give identical results -- also when the pools sized by a small batch overflow."""
    qs = _queries(submap, 6, seed=21)
    mk = lambda sub: capi.Scans([submap["angles"]] * len(sub), [s for s, _ in sub], [p for _, p in sub],
                                range_min=0.02, range_max=30.0)
    key = lambda o: (o.found, o.ix, o.iy, o.it, o.score)
    monkeypatch.setenv("LGS_BB_SYNC", "1")
    ref_batch = capi.BbBatch(ctx, **DEF)
    ref_batch.upload(mk(qs), [submap["pyr"]] * len(qs), 0.5)
    ref_batch.run()
    want = [key(o) for o in ref_batch.results()]
    want_levels = ref_batch.work()[0]
    monkeypatch.delenv("LGS_BB_SYNC")
    batch = capi.BbBatch(ctx, **DEF)
    batch.upload(mk(qs[:1]), [submap["pyr"]], 0.5)        # small batch: sizes the pools, leaves hints
    batch.run()
    assert [key(o) for o in batch.results()] == want[:1]
    batch.upload(mk(qs), [submap["pyr"]] * len(qs), 0.5)   # 6x the work: speculative run overflows -> redone
    batch.run()
    assert [key(o) for o in batch.results()] == want
    for _ in range(3):                                    # steady state: speculative runs validate
        batch.run()
        assert [key(o) for o in batch.results()] == want
        assert batch.work()[0] == want_levels


def test_bb_root_from_hit_points_with_a_few_near_edge_points(ctx, submap):
    """With at most 8 near-edge points the root level is scored straight from the hit points and only
    surviving (query, theta) rows get index rows; the flagged points still take the host's exact
    tables.  Widen the guard band just enough to flag a handful of points and compare with the
    reference; LGS_BB_TABLE=1 (full per-query table) must give the same answers."""
    import os
    from oracle import backend
    R = backend()
    qs = _queries(submap, 3, seed=33)
    scans = capi.Scans([submap["angles"]] * len(qs), [s for s, _ in qs], [p for _, p in qs],
                       range_min=0.02, range_max=30.0)
    refs = [R.bb_match(submap["refmap"], submap["angles"], scan, init, pyramid=submap["refpyr"],
                       thr=0.5, **_ref_kwargs(DEF)) for scan, init in qs]
    seen_small = False
    try:
        for eps in (1e-9, 3e-7, 6e-7, 1.2e-6, 2.4e-6):
            capi.set_edge_eps(eps)
            for table in (False, True):
                os.environ["LGS_BB_SLOTS"] = "1"       # one scan per query here: force the hit-point root path
                if table:
                    os.environ["LGS_BB_TABLE"] = "1"
                else:
                    os.environ.pop("LGS_BB_TABLE", None)
                batch = capi.BbBatch(ctx, **DEF)
                batch.upload(scans, [submap["pyr"]] * len(qs), 0.5)
                for rep in range(2):
                    batch.run()
                    outs = batch.results()
                    for out, ref in zip(outs, refs):
                        _same(out, ref)
                nfix = sum(o.n_fixups for o in outs)
                if not table and 1 <= nfix <= 8:
                    seen_small = True
                batch.close()
    finally:
        os.environ.pop("LGS_BB_TABLE", None)
        os.environ.pop("LGS_BB_SLOTS", None)
        capi.set_edge_eps(1e-9)
    assert seen_small, "no guard band produced between 1 and 8 near-edge points"
