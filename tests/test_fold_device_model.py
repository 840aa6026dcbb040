"""CPU model of the device's per-cell fold (csrc/lgs_integrate.cu, integ_touch / integ_fold) against the
reference's BinaryBayesGridCell::Update (grid_map/binary_bayes_grid_cell.hpp:75-119) applied touch by touch.

The device does not apply every touch: it packs a cell's ordered miss / hit sequence of one scan into a
32-bit record (raw bits up to 26 touches, three run lengths beyond) and the fold
  * skips, without arithmetic, touches that cannot change a value sitting on the probability clamp the
    observation pushes towards (the clamp makes the update idempotent there), and
  * leaves a run of equal touches at the first fixed point (nv == v).
Both shortcuts are arguments about the reference's arithmetic, checked here bit for bit on random and on
adversarial sequences, through the record formats."""
import numpy as np

LO, HI = 1e-3, 1.0 - 1e-3
RAW_MAX = 26


def _encode(seq):
    """Ordered touches of one (cell, scan) -> record (integ_touch_kernel, TouchSeq::record) or None."""
    n = len(seq)
    if n == 0:
        return 0
    if n <= RAW_MAX:
        return (n << 26) | sum(1 << k for k, h in enumerate(seq) if h)
    runs = []
    for h in seq:
        if runs and runs[-1][0] == h:
            runs[-1][1] += 1
        else:
            runs.append([h, 1])
    if len(runs) <= 3 and all(c < 1023 for _, c in runs):
        r = [c for _, c in runs] + [0, 0]
        return 0x80000000 | (int(runs[0][0]) << 30) | (r[0] << 20) | (r[1] << 10) | r[2]
    return None                                             # side buffer / re-derivation: raw order again


class Fold:
    """The fold thread's state machine for one cell."""

    def __init__(self, R, v, p_hit, p_miss):
        self.R, self.v, self.p = R, v, (p_miss, p_hit)
        self.miss_sat = R.bayes_update(LO, p_miss) == LO
        self.hit_sat = R.bayes_update(HI, p_hit) == HI
        self.computed = 0

    def _saturated(self, hit):
        return (self.hit_sat and self.v == HI) if hit else (self.miss_sat and self.v == LO)

    def touch(self, hit):
        if not self._saturated(hit):
            self.v = self.R.bayes_update(self.v, self.p[hit])
            self.computed += 1

    def run(self, hit, n):
        for _ in range(n):
            if self._saturated(hit):
                break
            nv = self.R.bayes_update(self.v, self.p[hit])
            self.computed += 1
            if nv == self.v:
                break                                       # fixed point: the rest of the run is a no-op
            self.v = nv

    def record(self, rec, seq):
        if rec is None:
            for h in seq:
                self.touch(h)
        elif rec >> 31:
            first = (rec >> 30) & 1
            self.run(first, (rec >> 20) & 1023)
            self.run(1 - first, (rec >> 10) & 1023)
            self.run(first, rec & 1023)
        else:
            n, bits = rec >> 26, rec & ((1 << 26) - 1)
            for k in range(n):
                self.touch((bits >> k) & 1)


def _sequences(rng):
    yield [0] * 700 + [1] * 300 + [0] * 5                                  # M^a H^b M^c, the near field
    yield [1] * 1000 + [0] * 1000
    yield [int(b) for b in rng.integers(0, 2, 26)]
    yield [int(b) for b in rng.integers(0, 2, 27)]                          # fits no record format
    for _ in range(60):
        kind = rng.integers(0, 4)
        if kind == 0:                                                       # short raw records
            yield [int(b) for b in rng.integers(0, 2, int(rng.integers(1, 27)))]
        elif kind == 1:                                                     # up to three long runs
            first = int(rng.integers(0, 2))
            lens = [int(rng.integers(1, 400)) for _ in range(int(rng.integers(1, 4)))]
            yield [first ^ (k & 1) for k, n in enumerate(lens) for _ in range(n)]
        elif kind == 2:                                                     # mostly misses with stray hits
            s = [0] * int(rng.integers(30, 300))
            for k in rng.integers(0, len(s), int(rng.integers(1, 5))):
                s[int(k)] = 1
            yield s
        else:
            yield []


def test_fold_with_skips_and_run_records_equals_touch_by_touch_updates():
    from oracle import backend
    R = backend()
    rng = np.random.default_rng(11)
    for p_hit, p_miss in ((0.6, 0.45), (0.62, 0.38), (0.9, 0.1), (0.5, 0.5), (0.55, 0.499)):
        skipped_any = False
        for start in (0.0, LO, HI, 0.3, 0.73):
            ref_v = start
            fold = Fold(R, start, p_hit, p_miss)
            touches = 0
            for seq in _sequences(rng):                                     # one record per "scan", in order
                for h in seq:
                    ref_v = R.bayes_update(ref_v, p_hit if h else p_miss)
                touches += len(seq)
                fold.record(_encode(seq), seq)
                assert np.float64(fold.v).view(np.int64) == np.float64(ref_v).view(np.int64)
            skipped_any |= fold.computed < touches
        assert skipped_any or p_hit == p_miss == 0.5
