"""CPU model of the device's correlative-matcher SELECTION against the reference's pruned loop.

The device sums every hypothesis of the window (csm_sweep_*), so what has to be argued is that the pruned
CPU loop (scan_matcher_real_time_correlative.cpp:88-115, :227-256) can be reproduced from the full tables
(csm_select_kernel, csrc/lgs_csm.cu):

  * per coarse block b the fine maximum F[b] with its FIRST-visited argmax (fx ascending, then fy);
  * if every coarse score bounds its block (C[b] >= F[b]) the pruned loop returns the first visit of the
    global fine maximum: a (score desc, visit asc) reduction;
  * otherwise (a negative coarse index reads 0 while fine cells are in the map) the CPU's
    sequence `if (C[b] > best && F[b] > best) best = F[b]` is replayed in visit order (theta, x, y).

The tables come from the reference's own ComputeScanIndices / ComputeScore (oracle score tables); the test
also records how much of the window the reference evaluates (what an exact coarse-to-fine sweep on the
device could save)."""
import numpy as np

from my_lidar_graph_slam_b200 import synth


def _select(fine, coarse, low_res, thr):
    """fine[t][y][x] (x, y from -win), coarse[t][y][x] at the same offsets -> (found, ix, iy, it, score,
    replayed, fine blocks the CPU evaluates)."""
    nt, nyw, nxw = fine.shape
    nbx, nby = nxw // low_res, nyw // low_res
    # blocks in CPU visit order: theta, then x, then y; within a block fx ascending, then fy
    F = np.empty((nt, nbx, nby))
    A = np.empty((nt, nbx, nby, 2), dtype=np.int64)
    C = np.empty((nt, nbx, nby))
    for bx in range(nbx):
        for by in range(nby):
            blk = fine[:, by * low_res:(by + 1) * low_res, bx * low_res:(bx + 1) * low_res]   # [t][fy][fx]
            vis = blk.transpose(0, 2, 1).reshape(nt, -1)                                      # visit order fx, fy
            k = vis.argmax(axis=1)                                                            # first maximum
            F[:, bx, by] = vis[np.arange(nt), k]
            A[:, bx, by, 0], A[:, bx, by, 1] = bx * low_res + k // low_res, by * low_res + k % low_res
            C[:, bx, by] = coarse[:, by * low_res, bx * low_res]
    bounded = bool((C >= F).all())
    if bounded:
        best = F.max()
        if not best > thr:
            return (0, None, None, None, thr, False, 0)
        t, bx, by = np.unravel_index(int(np.argmax(F == best)), F.shape)      # first in (t, bx, by) order
        evaluated = None
    else:
        best, t, bx, by = thr, None, None, None
        for tt in range(nt):
            for x in range(nbx):
                for y in range(nby):
                    if C[tt, x, y] > best and F[tt, x, y] > best:
                        best, t, bx, by = F[tt, x, y], tt, x, y
        if t is None:
            return (0, None, None, None, thr, True, 0)
    # blocks whose coarse score beats the running best when the CPU gets there
    run, evaluated = thr, 0
    for tt in range(nt):
        for x in range(nbx):
            for y in range(nby):
                if C[tt, x, y] > run:
                    evaluated += 1
                    run = max(run, F[tt, x, y])
    return (1, int(A[t, bx, by, 0]), int(A[t, bx, by, 1]), int(t), float(best), not bounded, evaluated)


def _check(R, refmap, pre, angles, ranges, init, params):
    ref = R.rtcsm_match(refmap, angles, ranges, init, pre=pre, **params)
    L = params["low_res"]
    nxw, nyw = ((2 * ref.winX) // L + 1) * L, ((2 * ref.winY) // L + 1) * L
    fine, _, _ = R.rtcsm_score_table(refmap, pre, False, L, params["scan_range_max"], init, angles, ranges,
                                     ref.stepT, ref.winT, -ref.winX, nxw, -ref.winY, nyw)
    coarse, _, _ = R.rtcsm_score_table(refmap, pre, True, L, params["scan_range_max"], init, angles, ranges,
                                       ref.stepT, ref.winT, -ref.winX, nxw, -ref.winY, nyw)
    thr = float(np.finfo(np.float64).tiny) * len(ranges)        # the 1-argument overload (:45-47)
    found, ix, iy, it, score, replayed, evaluated = _select(fine, coarse, L, thr)
    assert found == ref.found
    if found:
        assert (ix - ref.winX, iy - ref.winY, it - ref.winT) == (ref.ix, ref.iy, ref.it)
        assert score == ref.score
    blocks = fine.shape[0] * (nxw // L) * (nyw // L)
    return replayed, evaluated / blocks, float(fine.max()) != ref.score


def test_selection_from_full_tables_equals_the_pruned_loop():
    from oracle import backend
    from scenes import room_scene
    R = backend()
    world, angles, traj, builder = room_scene(seed=1, n_beams=361)
    refmap = builder.latest_map()
    pre = refmap.precompute(5)
    params = dict(low_res=5, range_x=1.0, range_y=1.0, range_theta=0.6, scan_range_max=5.7296)
    rng = np.random.default_rng(3)
    fracs = []
    for k in range(4):
        true = traj[10 + k]
        scan = synth.make_scan(world, true, angles, np.random.default_rng(100 + k))
        init = true + np.array([rng.uniform(-0.3, 0.3), rng.uniform(-0.3, 0.3), rng.uniform(-0.2, 0.2)])
        _, frac, _ = _check(R, refmap, pre, angles, scan, init, params)
        fracs.append(frac)
    # the reference's loop really skips most fine blocks -- and still evaluates a good part of them
    assert 0.05 < min(fracs) and max(fracs) < 0.9


def test_selection_replays_the_cpu_order_where_coarse_scores_are_no_bounds():
    """H12: scans hanging over the lower-left map edge; the reference's answer differs from the exhaustive
    optimum in some of these scenes and the replay returns the reference's."""
    from oracle import backend
    R = backend()
    rng = np.random.default_rng(7)
    ny, nx = 128, 128
    dense = np.where(rng.random((ny, nx)) < 0.25, rng.uniform(0.05, 0.95, (ny, nx)), 0.0)
    dense[:12, :] = rng.uniform(0.5, 0.99, (12, nx))
    dense[:, :12] = rng.uniform(0.5, 0.99, (ny, 12))
    refmap = R.RefMap.from_dense(dense, -1.0, -2.0)
    pre = refmap.precompute(5)
    params = dict(low_res=5, range_x=1.0, range_y=1.0, range_theta=0.2, scan_range_max=20.0)
    angles = synth.beam_angles(181, 180.0)
    replayed = differs = 0
    for k in range(10):
        ranges = rng.uniform(0.1, 0.5, angles.shape)
        init = np.array([-1.0 + rng.uniform(0.0, 0.3), -2.0 + rng.uniform(0.5, 3.0), np.pi + rng.uniform(-0.3, 0.3)])
        r, _, d = _check(R, refmap, pre, angles, ranges, init, params)
        replayed += r
        differs += d
    assert replayed > 0 and differs > 0
