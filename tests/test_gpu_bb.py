"""GPU parity: branch-and-bound matcher through the C ABI vs the reference's ScanMatcherBranchBound.

Winning (ix, iy, itheta) and found flag bit-exact; score bit-identical (bar: 1e-5 relative)."""
import numpy as np
import pytest

from my_lidar_graph_slam_b200 import capi, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["device", "exact"])
def bb_run_path(request, ctx):
    """Every test runs twice: through the device-only run (ONE persistent kernel: fixed-point cells,
    near-edge points decided on the device) and through the level-synchronous exact path (full index
    table, near-edge points from the host)."""
    ctx.set_option("bb_sync", 1 if request.param == "exact" else 0)
    yield request.param
    ctx.set_option("bb_sync", 0)

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
    """Widen the edge guard band so ~5 % of the projected points take the exact evaluation (on the
    device: interval evaluation of the CPU's expression; exact path: the host's glibc tables)."""
    from oracle import backend
    R = backend()
    qs = _queries(submap, 2, seed=21)
    ctx.set_edge_eps(0.025)
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
        ctx.set_edge_eps(1e-9)


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


def test_device_only_runs_match_the_exact_path_and_survive_pool_overflow(ctx, submap, bb_run_path):
    """A device-only run never talks to the host; its node pools are sized from earlier runs.  A batch
    that outgrows them (low threshold: almost the whole tree survives) is detected on the device and
    repeated on the exact path, which also re-sizes the pools for the following device-only runs."""
    qs = _queries(submap, 4, seed=21)
    params = dict(DEF, node_height_max=4, range_x=1.0, range_y=1.0, range_theta=0.3)
    pyr = capi.Pyramid(ctx, submap["grid"], 4)
    mk = lambda sub: capi.Scans([submap["angles"]] * len(sub), [s for s, _ in sub], [p for _, p in sub],
                                range_min=0.02, range_max=30.0)
    key = lambda o: (o.found, o.ix, o.iy, o.it, o.score)
    ctx.set_option("bb_sync", 1)
    ref_batch = capi.BbBatch(ctx, **params)
    want, want_levels = {}, {}
    for thr in (0.5, 0.02):
        ref_batch.upload(mk(qs), [pyr] * len(qs), thr)
        ref_batch.run()
        want[thr] = [key(o) for o in ref_batch.results()]
        want_levels[thr] = ref_batch.work()[0]
    ctx.set_option("bb_sync", 1 if bb_run_path == "exact" else 0)
    assert sum(want_levels[0.02]) > 4 * sum(want_levels[0.5])
    batch = capi.BbBatch(ctx, **params)
    batch.upload(mk(qs), [pyr] * len(qs), 0.5)
    for _ in range(2):
        batch.run()
        assert [key(o) for o in batch.results()] == want[0.5]
        assert batch.work()[0] == want_levels[0.5]
    dev0, exact0 = batch.path()
    batch.upload(mk(qs), [pyr] * len(qs), 0.02)            # the pools sized so far overflow
    batch.run()
    assert [key(o) for o in batch.results()] == want[0.02]
    assert batch.work()[0] == want_levels[0.02]
    dev1, exact1 = batch.path()
    if bb_run_path != "exact":
        assert dev0 == 2 and exact0 == 0
        assert dev1 == 3 and exact1 == 1                  # the overflowed run was repeated exactly
    for _ in range(2):                                    # steady state: the grown pools hold the tree
        batch.run()
        assert [key(o) for o in batch.results()] == want[0.02]
    if bb_run_path != "exact":
        assert batch.path() == (5, 1)


def test_undecided_near_edge_points_fall_back_to_the_exact_path(ctx, submap, bb_run_path):
    """The device decides a near-edge point by evaluating the CPU's expression at both ends of an
    interval around its own cos / sin.  With an absurdly wide interval ("bb_resolve_ulps") and a wide
    guard band the ends disagree for many points: the run must report them, be repeated on the exact
    path, and still return the reference's answer."""
    from oracle import backend
    R = backend()
    (scan, init), = _queries(submap, 1, seed=44)
    ref = R.bb_match(submap["refmap"], submap["angles"], scan, init, pyramid=submap["refpyr"],
                     thr=0.6, **_ref_kwargs(DEF))
    try:
        ctx.set_edge_eps(1e-4)
        ctx.set_option("bb_resolve_ulps", 2_000_000_000)
        batch = capi.BbBatch(ctx, **DEF)
        batch.upload(capi.Scans([submap["angles"]], [scan], [init], range_min=0.02, range_max=30.0),
                     [submap["pyr"]], 0.6)
        batch.run()
        (out,) = batch.results()
        _same(out, ref)
        if bb_run_path != "exact":
            assert batch.path() == (1, 1) and out.reserved == 1
    finally:
        ctx.set_edge_eps(1e-9)
        ctx.set_option("bb_resolve_ulps", 8)


def test_a_few_near_edge_points_are_decided_identically_on_both_paths(ctx, submap):
    """Widen the guard band just enough to flag a handful of points and compare with the reference;
    the device-only run (on-device interval evaluation) and the exact path (host tables) must agree."""
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
            ctx.set_edge_eps(eps)
            batch = capi.BbBatch(ctx, **DEF)
            batch.upload(scans, [submap["pyr"]] * len(qs), 0.5)
            for rep in range(2):
                batch.run()
                outs = batch.results()
                for out, ref in zip(outs, refs):
                    _same(out, ref)
            nfix = sum(o.n_fixups for o in outs)
            if 1 <= nfix <= 16:
                seen_small = True
            batch.close()
    finally:
        ctx.set_edge_eps(1e-9)
    assert seen_small, "no guard band produced between 1 and 16 near-edge points"


def test_records_and_device_sink(ctx, submap):
    """The finalize phase writes one 32-byte record per pair: into the batch's own buffer and, when a
    sink is set, at sink[first_slot + q] -- here a second batch's record buffer stands in for the
    peer-mapped gather buffer of another GPU."""
    qs = _queries(submap, 5, seed=8)
    scans = capi.Scans([submap["angles"]] * len(qs), [s for s, _ in qs], [p for _, p in qs],
                       range_min=0.02, range_max=30.0)
    batch = capi.BbBatch(ctx, **DEF)
    ids = np.array([70, 10, 40, 20, 90], dtype=np.int64)
    batch.set_record_ids(ids)
    batch.upload(scans, [submap["pyr"]] * len(qs), 0.6)
    batch.run()
    outs = batch.results()
    rec = batch.records()
    assert rec["submap"].tolist() == ids.tolist()
    for o, r in zip(outs, rec):
        assert (o.found, o.ix, o.iy, o.it, o.score) == (int(r["found"]), int(r["ix"]), int(r["iy"]), int(r["it"]), float(r["score"]))
    # sink: slots 3.. of a separate device buffer, followed by the run's status record
    buf = capi.device_alloc(ctx, 12 * 32)
    batch.set_record_sink(buf, 3)
    batch.run()
    batch.settle()
    got = capi.download_records(ctx, buf, 12)
    assert got[3:8].tobytes() == rec.tobytes()
    assert int(got[8]["submap"]) == -1 and int(got[8]["found"]) == 1          # status: the run stands
    assert not got[:3].tobytes().strip(b"\0") and not got[9:].tobytes().strip(b"\0")
    batch.set_record_sink(0)
    batch.close()
    capi.device_free(ctx, buf)


def test_per_query_node_counts(ctx, submap, bb_run_path):
    """With "bb_count_nodes" on, a device-only run reports the nodes it scored below the root level per
    query (what a cost-aware placement of submaps on devices balances): they add up to the batch's
    level totals, n_scored becomes per query, and the results do not change."""
    if bb_run_path == "exact":
        return                                                        # node counts belong to the device-only run
    qs = _queries(submap, 6, seed=11)
    scans = capi.Scans([submap["angles"]] * len(qs), [s for s, _ in qs], [p for _, p in qs],
                       range_min=0.02, range_max=30.0)
    batch = capi.BbBatch(ctx, **DEF)
    batch.upload(scans, [submap["pyr"]] * len(qs), 0.55)
    batch.run()
    batch.results_array()                                             # the first run sizes the node pools
    exact0 = batch.path()[1]
    batch.run()
    plain = batch.results_array().copy()
    with pytest.raises(capi.LgsError):
        batch.query_nodes(len(qs))
    ctx.set_option("bb_count_nodes", 1)
    try:
        batch.run()
        res = batch.results_array().copy()
        nodes = batch.query_nodes(len(qs))
    finally:
        ctx.set_option("bb_count_nodes", 0)
    levels, _ = batch.work()
    assert batch.path()[1] == exact0                                  # both runs stayed on the device
    assert int(nodes.sum()) == int(sum(levels[:-1]))
    assert int(res["n_scored"].sum()) == int(sum(levels))
    assert len(set(plain["n_scored"].tolist())) == 1 and int(plain["n_scored"][0]) == int(sum(levels))
    for f in ("found", "ix", "iy", "it", "score"):
        assert np.array_equal(plain[f], res[f])
    batch.close()
