"""GPU parity, randomised: many small seeded scenes with random sensor models, map sizes, search
windows, thresholds and filter probabilities -- integration, the three matchers, the matchers' tail and
the pyramid each time, every result compared bit for bit with the reference's own classes."""
import numpy as np
import pytest

from my_lidar_graph_slam_b200 import capi, synth

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a).view(np.int64)


@pytest.mark.parametrize("seed", range(24))
def test_random_scene_end_to_end(ctx, seed):
    from oracle import backend
    R = backend()
    rng = np.random.default_rng(1000 + seed)
    n_beams = int(rng.choice([91, 181, 361, 541, 721, 1081]))
    fov = float(rng.choice([120.0, 180.0, 240.0, 270.0, 359.0]))
    size = float(rng.uniform(10.0, 26.0))
    world = synth.World(size, size, int(rng.integers(3, 12)), seed=seed + 50)
    angles = synth.beam_angles(n_beams, fov)
    n_map = int(rng.integers(4, 14))
    traj = synth.trajectory(world, n_map + 3, step=float(rng.uniform(0.1, 0.4)), seed=seed + 7)
    noise = np.random.default_rng(seed)
    p_hit, p_miss = float(rng.uniform(0.55, 0.9)), float(rng.uniform(0.1, 0.48))
    rmax = float(rng.uniform(6.0, 25.0))
    scans = [synth.make_scan(world, p, angles, noise) for p in traj]

    # ---- integration: ConstructMapFromScans geometry (tight Resize) + one device batch
    builder = R.RefBuilder(n_latest=n_map, rmax=rmax, p_hit=p_hit, p_miss=p_miss)
    for p, r in zip(traj[:n_map], scans[:n_map]):
        builder.append_scan(p, angles, r)
    refmap = builder.latest_map()
    nx, ny, mx, my, res = refmap.geometry()
    hp = [capi.scan_hit_points(p, angles, r, 0.02, min(rmax, 30.0)) for p, r in zip(traj[:n_map], scans[:n_map])]
    grid = capi.Grid(ctx, nx, ny, mx, my, res, apron=64)
    capi.integrate_scans(ctx, grid, traj[:n_map, :2], [h for h, _ in hp], p_hit, p_miss)
    assert np.array_equal(_bits(grid.download()), _bits(refmap.dense())), "integration"

    # ---- pyramid
    H = int(rng.integers(3, 7))
    pyr = capi.Pyramid(ctx, grid, H)
    refpyr = refmap.pyramid(H)
    for lvl in range(H + 1):
        assert np.array_equal(_bits(pyr.download(lvl)), _bits(refpyr[lvl].dense())), f"pyramid level {lvl}"

    # ---- correlative matcher
    low = int(rng.integers(1, 9))
    cp = dict(low_res=low, range_x=float(rng.uniform(0.2, 1.2)), range_y=float(rng.uniform(0.2, 1.2)),
              range_theta=float(rng.uniform(0.1, 0.8)), scan_range_max=float(rng.uniform(4.0, 20.0)))
    thr = None if rng.random() < 0.4 else float(rng.uniform(0.2, 0.8))
    coarse = grid.precompute(low)
    pre = refmap.precompute(low)
    assert np.array_equal(_bits(coarse.download()), _bits(pre.dense())), "coarse map"
    batch = capi.RtcsmBatch(ctx, **cp)
    for k in range(2):
        true = traj[n_map + k]
        init = true + np.array([rng.uniform(-0.25, 0.25), rng.uniform(-0.25, 0.25), rng.uniform(-0.15, 0.15)])
        ref = R.rtcsm_match(refmap, angles, scans[n_map + k], init, pre=pre, thr=thr, **cp)
        batch.upload(grid, capi.Scans([angles], [scans[n_map + k]], [init]), None if thr is None else [thr])
        batch.run(grid, coarse)
        (out,) = batch.results(grid, coarse)
        assert (out.found, out.ix, out.iy, out.it) == (ref.found, ref.ix, ref.iy, ref.it), ("rtcsm", k)
        if ref.found:
            assert out.score == ref.score

    # ---- branch-and-bound matcher
    bp = dict(node_height_max=H, range_x=float(rng.uniform(0.5, 2.5)), range_y=float(rng.uniform(0.5, 2.5)),
              range_theta=float(rng.uniform(0.2, 0.9)), scan_range_max=float(rng.uniform(5.0, 20.0)),
              score_range_min=0.01, score_range_max=float(rng.uniform(5.0, 25.0)))
    bthr = float(rng.uniform(0.25, 0.7))
    bb = capi.BbBatch(ctx, **bp)
    qs = []
    for k in range(2):
        true = traj[n_map + k]
        qs.append((scans[n_map + k], true + np.array([rng.uniform(-0.5, 0.5), rng.uniform(-0.5, 0.5),
                                                       rng.uniform(-0.2, 0.2)])))
    bb.upload(capi.Scans([angles] * 2, [s for s, _ in qs], [p for _, p in qs], range_min=0.02, range_max=30.0),
              [pyr, pyr], bthr)
    refs = [R.bb_match(refmap, angles, scan, init, pyramid=refpyr, thr=bthr, height_max=H,
                       range_x=bp["range_x"], range_y=bp["range_y"], range_theta=bp["range_theta"],
                       scan_range_max=bp["scan_range_max"], score_range_min=0.01,
                       score_range_max=bp["score_range_max"]) for scan, init in qs]
    for rep in range(3):                      # device-only runs (one persistent kernel); the last one is forced
        ctx.set_option("bb_sync", 1 if rep == 2 else 0)      # through the level-synchronous exact path
        try:
            bb.run()
        finally:
            ctx.set_option("bb_sync", 0)
        for k, (out, ref) in enumerate(zip(bb.results(), refs)):
            assert (out.found, out.ix, out.iy, out.it) == (ref.found, ref.ix, ref.iy, ref.it), ("bb", k, rep)
            if ref.found:
                assert out.score == ref.score

    # ---- the matchers' tail at the correlative winner, random cost-function parameters
    cs = (float(rng.uniform(0.0, 0.3)), float(rng.uniform(5.0, 25.0)), float(rng.uniform(0.03, 0.15)),
          float(rng.uniform(0.05, 0.5)), float(rng.integers(0, 4)), float(rng.uniform(0.05, 2.0)),
          float(rng.uniform(0.03, 1.0)))
    tail_scans = capi.Scans([angles] * 2, [s for s, _ in qs], [p for _, p in qs], range_min=0.02, range_max=30.0)
    nc, cov, _ = capi.cost_tail(ctx, grid, tail_scans, [p for _, p in qs], cost=cs)
    for k, (scan, pose) in enumerate(qs):
        wn, _, wc = R.host_tail(refmap, pose, angles, scan, cost=cs)
        assert nc[k] == wn and np.array_equal(cov[k], wc), ("tail", k)

    # ---- exhaustive grid search with free (non-resolution) steps, small window
    gp = dict(range_x=float(rng.uniform(0.1, 0.5)), range_y=float(rng.uniform(0.1, 0.5)),
              range_theta=float(rng.uniform(0.02, 0.12)), step_x=float(rng.uniform(0.02, 0.12)),
              step_y=float(rng.uniform(0.02, 0.12)), step_theta=float(rng.uniform(0.004, 0.03)),
              score_range_min=0.01, score_range_max=float(rng.uniform(5.0, 25.0)))
    gthr = float(rng.uniform(0.2, 0.7))
    gout = capi.gs_match(ctx, tail_scans, [grid, grid], norm_threshold=gthr, **gp)
    for k, (scan, pose) in enumerate(qs):
        ref = R.gs_match(refmap, angles, scan, pose, thr=gthr, **gp)
        got = gout[k]
        assert (got.found, got.ix, got.iy, got.it, got.win_x, got.win_y, got.win_t) == \
            (ref.found, ref.ix, ref.iy, ref.it, ref.winX, ref.winY, ref.winT), ("gs", k)
        if ref.found:
            assert got.score == ref.score
