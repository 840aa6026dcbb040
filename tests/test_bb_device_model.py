"""CPU model of the device's branch-and-bound ALGORITHM against the reference's depth-first search.

The GPU tests (-m gpu) compare the kernels with the reference; this file checks, without a GPU, the
reasoning the persistent kernel rests on (csrc/lgs_bb_run.cu, csrc/lgs_bb.cu), restated in
numpy with the reference's own arithmetic:

  * breadth first over the superset "every ancestor beats the STATIC threshold" with stored scores;
  * LIFO visit ranks (root pop order x 4^H + child digits; children pop (x+w, y+w), (x, y+w), (x+w, y), (x, y));
  * winner = (score desc, rank asc) over the leaves; accepted if every ancestor scores >= the winner,
    otherwise the CPU's stack is replayed over the stored scores;
  * early rejection: a node's beam-order sum stops after a round of 16 usable beams once
    partial + remaining beams <= threshold, and the partial sum is what gets stored.

Node scores use ScorePixelAccurate's expressions (score_function_pixel_accurate.cpp:19-76) in the CPU's
operation order with libm cos / sin, summed one beam at a time (np.cumsum is sequential)."""
import math

import numpy as np
import pytest

from my_lidar_graph_slam_b200 import synth

DEF = dict(height_max=6, range_x=2.0, range_y=2.0, range_theta=1.0, scan_range_max=20.0,
           score_range_min=0.01, score_range_max=20.0)


class Model:
    """One (scan, submap) query the way the device runs it."""

    def __init__(self, levels, geom, ref, angles, ranges, params, thr_norm, early_reject=True,
                 scan_min_range=0.02, scan_max_range=30.0):
        self.levels, (self.nx, self.ny, self.min_x, self.min_y, self.res) = levels, geom
        self.H = params["height_max"]
        lo = max(params["score_range_min"], scan_min_range)
        hi = min(params["score_range_max"], scan_max_range)
        keep = ~((ranges >= hi) | (ranges <= lo))                      # :27-41 range filter
        self.r, self.a = np.asarray(ranges)[keep], np.asarray(angles)[keep]
        self.n_use = int(keep.sum())
        self.sensor = np.array(ref.sensorPose[:3])
        self.step_t, self.win = ref.stepT, (ref.winX, ref.winY, ref.winT)
        # static threshold: normalised threshold x ALL beams (scan_matcher_branch_bound.cpp:75-76)
        self.thr = (np.finfo(np.float64).tiny if thr_norm is None else thr_norm) * float(len(ranges))
        self.early = early_reject
        self.trig = {}
        self.skipped = 0

    def _trig(self, t):
        if t not in self.trig:
            th = self.sensor[2] + t * self.step_t                       # :96-99, sensor_data.hpp:171-172
            self.trig[t] = (np.array([math.cos(th + a) for a in self.a]), np.array([math.sin(th + a) for a in self.a]))
        return self.trig[t]

    def score(self, x, y, t, h):
        """The value the device stores for node (x, y, t) of level h."""
        c, s = self._trig(t)
        hx = (self.sensor[0] + x * self.res) + self.r * c
        hy = (self.sensor[1] + y * self.res) + self.r * s
        ix = np.floor((hx - self.min_x) / self.res).astype(np.int64)    # grid_map.hpp:779-790
        iy = np.floor((hy - self.min_y) / self.res).astype(np.int64)
        ok = (ix >= 0) & (ix < self.nx) & (iy >= 0) & (iy < self.ny)
        v = np.where(ok, self.levels[h][np.clip(iy, 0, self.ny - 1), np.clip(ix, 0, self.nx - 1)], 0.0)
        pre = np.cumsum(v)                                              # beam order, one add at a time
        if self.n_use == 0:
            return 0.0
        if self.early:
            lim = self.thr - 1e-6
            for k in range(16, self.n_use - self.n_use % 16 + 1, 16):   # after every full round of 16 beams
                if pre[k - 1] + float(self.n_use - k) <= lim:
                    self.skipped += self.n_use - k
                    return float(pre[k - 1])
        return float(pre[-1])

    def run(self):
        H, (wx, wy, wt) = self.H, self.win
        W = 1 << H
        roots = [(x, y, t) for x in range(-wx, wx + 1, W) for y in range(-wy, wy + 1, W) for t in range(-wt, wt + 1)]
        n_roots = len(roots)
        # level pools: node = (x, y, t, rank, parent index); LIFO pops reverse the push order
        pools = {H: [(x, y, t, n_roots - 1 - p, -1) for p, (x, y, t) in enumerate(roots)]}
        scores, children = {}, {}
        for h in range(H, -1, -1):
            scores[h] = [self.score(x, y, t, h) for (x, y, t, _, _) in pools[h]]
            if h == 0:
                break
            nxt, children[h] = [], {}
            w = 1 << (h - 1)
            for k, ((x, y, t, rank, _), sc) in enumerate(zip(pools[h], scores[h])):
                if sc > self.thr:                                       # :108 with scoreMax >= threshold
                    children[h][k] = len(nxt)
                    for cdig, (dx, dy) in enumerate(((w, w), (0, w), (w, 0), (0, 0))):   # pop order :134-137
                        nxt.append((x + dx, y + dy, t, rank * 4 + cdig, k))
            pools[h - 1] = nxt
        leaves = [(sc, rank, k) for k, ((_, _, _, rank, _), sc) in enumerate(zip(pools[0], scores[0])) if sc > self.thr]
        if not leaves:
            return (0, 0, 0, 0, self.thr, False)
        best = max(sc for sc, _, _ in leaves)
        _, _, k = min((rank, kk, kk) for sc, rank, kk in leaves if sc == best)
        ok, idx = True, k
        for h in range(0, H):
            idx = pools[h][idx][4]
            ok &= scores[h + 1][idx] >= best
        if ok:
            x, y, t, _, _ = pools[0][k]
            return (1, x, y, t, best, False)
        # CPU-order replay over the stored superset scores
        score_max, win = self.thr, None
        stack = [(H, p) for p in range(n_roots)]                        # popped from the end = rank order
        while stack:
            h, k = stack.pop()
            sc = scores[h][k]
            if sc <= score_max:
                continue
            if h == 0:
                score_max, win = sc, pools[0][k]
                continue
            base = children[h][k]                                       # sc > best >= thr => expanded
            stack.extend((h - 1, base + c) for c in (3, 2, 1, 0))       # so that child 0 pops first
        if win is None:
            return (0, 0, 0, 0, self.thr, True)
        return (1, win[0], win[1], win[2], score_max, True)


def _check(R, refmap, refpyr, angles, ranges, init, params, thr, early=True):
    ref = R.bb_match(refmap, angles, ranges, init, pyramid=refpyr,
                     thr=thr if thr is not None else float(np.finfo(np.float64).tiny), **params)
    nx, ny, mx, my, res = refmap.geometry()
    levels = [np.asarray(p.dense()) for p in refpyr]
    m = Model(levels, (nx, ny, mx, my, res), ref, angles, ranges, params, thr, early_reject=early)
    found, x, y, t, score, replayed = m.run()
    assert (found, x, y, t) == (ref.found, ref.ix, ref.iy, ref.it)
    if ref.found:
        assert score == ref.score
    return ref.found, replayed, m.skipped


@pytest.fixture(scope="module")
def scene():
    from oracle import backend
    R = backend()
    world = synth.RoomsWorld(24.0, 4.0, seed=3)
    angles = synth.beam_angles(361, 240.0)
    traj = synth.trajectory(world, 40, step=0.25, seed=3)
    noise = np.random.default_rng(2)
    builder = R.RefBuilder()
    for p in traj[:30]:
        builder.append_scan(p, angles, synth.make_scan(world, p, angles, noise))
    refmap = builder.local_map(0)
    return dict(R=R, world=world, angles=angles, traj=traj, refmap=refmap)


@pytest.mark.parametrize("params,thr", [
    (dict(DEF, height_max=4, range_x=1.0, range_y=1.0, range_theta=0.3), 0.5),
    (dict(DEF, height_max=3, range_x=0.6, range_y=0.8, range_theta=0.2), 0.3),
    (dict(DEF, height_max=5, range_x=1.0, range_y=1.0, range_theta=0.2), 0.95),    # nothing found
    (dict(DEF, height_max=2, range_x=0.3, range_y=0.3, range_theta=0.04), None),   # DBL_MIN: the whole tree
])
def test_breadth_first_superset_with_early_rejection_returns_the_references_answer(scene, params, thr):
    R, rng = scene["R"], np.random.default_rng(5)
    refpyr = scene["refmap"].pyramid(params["height_max"])
    found = skipped = 0
    for k in range(5):
        true = scene["traj"][int(rng.integers(4, 28))]
        scan = synth.make_scan(scene["world"], true, scene["angles"], np.random.default_rng(300 + k))
        init = true + np.array([rng.uniform(-0.3, 0.3), rng.uniform(-0.3, 0.3), rng.uniform(-0.1, 0.1)])
        f, _, sk = _check(R, scene["refmap"], refpyr, scene["angles"], scan, init, params, thr)
        found += f
        skipped += sk
        # and without early rejection (what "bb_early_reject" = 0 runs)
        _check(R, scene["refmap"], refpyr, scene["angles"], scan, init, params, thr, early=False)
    if thr is not None and thr < 0.9:
        assert found >= 3 and skipped > 0


def test_low_edge_overhang_needs_the_replay_and_gets_it_right():
    """H12: window indices straddling zero make the win-max values non-bounds (the reference reads 0 for a
    negative index although cells of that window are in the map): the ancestor check must notice and the
    replay over the stored scores -- partial sums included -- must return the reference's answer."""
    from oracle import backend
    R = backend()
    rng = np.random.default_rng(17)
    ny, nx = 128, 192
    dense = np.where(rng.random((ny, nx)) < 0.25, rng.uniform(0.05, 0.95, (ny, nx)), 0.0)
    dense[:16, :] = rng.uniform(0.5, 0.99, (16, nx))
    dense[:, :16] = rng.uniform(0.5, 0.99, (ny, 16))
    refmap = R.RefMap.from_dense(dense, -1.0, -2.0)
    params = dict(DEF, height_max=4, range_x=1.0, range_y=1.0, range_theta=0.2)
    refpyr = refmap.pyramid(4)
    angles = synth.beam_angles(181, 180.0)
    replays = 0
    for k in range(12):
        ranges = rng.uniform(0.1, 0.6, angles.shape)
        init = np.array([-1.0 + rng.uniform(0.0, 0.4), -2.0 + rng.uniform(0.3, 3.0), np.pi + rng.uniform(-0.3, 0.3)])
        thr = float(rng.uniform(0.2, 0.5))
        _, replayed, _ = _check(R, refmap, refpyr, angles, ranges, init, params, thr)
        replays += replayed
    assert replays > 0
