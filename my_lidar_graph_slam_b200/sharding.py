"""Multi-GPU partitioning of loop-detection work.

(scan, submap) branch-and-bound queries are independent, so submap i (with its device-resident
pyramid) lives on rank i mod world and every rank searches only its own pairs.  The one exchange
step is an all-gather of the fixed-size result records (found, ix, iy, itheta, score, id), after
which every rank holds the full result list in global order and can pick the best candidate.

Two drivers, both inside the C ABI (csrc/lgs_group.cu):
  * one process, all devices: capi.Group / capi.GroupBb (peer stores into the root's gather buffer);
  * one process per device (torchrun): RecordExchange below -- the batch's finalize phase writes the
    rank's records straight into its slice of the receive buffer and lgs_comm_all_gather_records runs
    the NCCL all-gather in place on the context stream; nothing touches the host until the gathered
    buffer is downloaded once.
torch.distributed is plumbing only (rendezvous of the 128-byte NCCL id; gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np

RECORD = np.dtype([("found", np.int32), ("ix", np.int32), ("iy", np.int32), ("it", np.int32),
                   ("score", np.float64), ("submap", np.int64)])   # 32 bytes = lgs_loop_record


def owned(n_items: int, rank: int, world: int) -> np.ndarray:
    """Global indices owned by `rank` under round-robin placement."""
    return np.arange(rank, n_items, world, dtype=np.int64)


def balanced_placement(weights, world: int):
    """Cost-aware placement of items (submaps) on `world` ranks: longest-processing-time-first greedy on
    measured weights (nodes the branch-and-bound matcher scored against each submap in earlier
    queries).  A few submaps -- the true loop candidates, whose search goes deep -- carry most of the
    work, so round-robin placement leaves the ranks that happen to hold two of them ~30 % above the
    mean; LPT is within 4/3 of the optimal makespan.  Deterministic (ties: lower index first, lower
    rank first), so every rank computes the same placement from the same all-reduced weights.
    Returns one ascending index array per rank."""
    w = np.asarray(weights, dtype=np.float64)
    order = np.lexsort((np.arange(len(w)), -w))              # weight descending, index ascending
    load = np.zeros(world, dtype=np.float64)
    count = np.zeros(world, dtype=np.int64)
    place = [[] for _ in range(world)]
    for i in order:
        r = int(np.lexsort((np.arange(world), count, load))[0])   # least load, then fewest items, then lowest rank
        place[r].append(int(i))
        load[r] += w[i]
        count[r] += 1
    return [np.asarray(sorted(p), dtype=np.int64) for p in place]


def round_robin_placement(n_items: int, world: int):
    """owned() for every rank, in the form balanced_placement returns."""
    return [owned(n_items, r, world) for r in range(world)]


def slots_per_rank(n_items: int, world: int) -> int:
    """Record slots every rank contributes to the all-gather: its (padded) share + one status record."""
    return (n_items + world - 1) // world + 1


def _segments(counts):
    """Normalise per-rank record counts: an int (one batch) or a list (one entry per sub-batch)."""
    return [[int(c)] if np.isscalar(c) else [int(x) for x in c] for c in counts]


def assemble(gathered: np.ndarray, counts, per: int, n_items: int):
    """Gathered buffer (world x per records) -> (records in global order, all runs valid?).  Rank r's
    slice holds, for each of its sub-batches in turn, `counts[r][k]` records followed by that run's
    status record (lgs_bb_batch_set_record_sink)."""
    out = np.zeros(n_items, dtype=RECORD)
    ok = True
    for r, seg in enumerate(_segments(counts)):
        pos = r * per
        for cnt in seg:
            sl = gathered[pos:pos + cnt]
            out[sl["submap"]] = sl
            ok &= int(gathered[pos + cnt]["found"]) == 1
            pos += cnt + 1
    return out, bool(ok)


def pack_array(results: np.ndarray, global_ids) -> np.ndarray:
    """Records from a structured result array (host-side path of the gloo tests)."""
    rec = np.zeros(len(global_ids), dtype=RECORD)
    for f in ("found", "ix", "iy", "it", "score"):
        rec[f] = results[f]
    rec["submap"] = global_ids
    return rec


def all_gather_host(local: np.ndarray, counts, n_items: int, rank: int, world: int):
    """The exchange of RecordExchange with host tensors over torch.distributed (gloo): used by the CPU
    tests of the partitioning / assembly logic, never by a GPU run.  counts[r] = records of rank r
    (every rank can compute all of them: ownership is a pure function of the indices)."""
    per = max(counts) + 1
    assert len(local) == counts[rank]
    if world == 1:
        out = np.zeros(n_items, dtype=RECORD)
        out[local["submap"]] = local
        return out
    import torch
    import torch.distributed as dist
    buf = np.zeros(per, dtype=RECORD)
    buf[:len(local)] = local
    buf[len(local)]["found"] = 1
    buf[len(local)]["submap"] = -1
    send = torch.from_numpy(buf.view(np.uint8).copy())
    recv = torch.empty(world * send.numel(), dtype=torch.uint8)
    dist.all_gather_into_tensor(recv, send)
    out, ok = assemble(recv.numpy().view(RECORD), counts, per, n_items)
    assert ok
    return out


def make_comm(ctx, rank: int, world: int):
    """lgs_comm over the ranks of an initialised torch.distributed job (the 128-byte NCCL id travels
    through broadcast_object_list; the data path never touches torch)."""
    from . import capi
    if world == 1:
        return None
    import torch.distributed as dist
    box = [capi.Comm.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    return capi.Comm(ctx, world, rank, box[0])


class RecordExchange:
    """Device-resident exchange of loop records between one-process-per-GPU ranks.

    counts[r] = number of records of rank r (an int), or a list with one entry per device sub-batch of
    rank r; every rank computes all of them, ownership being a pure function of the indices."""

    def __init__(self, ctx, comm, n_items: int, rank: int, world: int, counts=None):
        from . import capi
        self.ctx, self.comm, self.n_items, self.rank, self.world = ctx, comm, int(n_items), rank, world
        if counts is None:
            counts = [len(owned(n_items, r, world)) for r in range(world)]
        self.counts = _segments(counts)
        self.per = max(sum(c + 1 for c in seg) for seg in self.counts)
        self.buf = capi.device_alloc(ctx, world * self.per * RECORD.itemsize)
        self.first = np.concatenate([[0], np.cumsum([c + 1 for c in self.counts[rank]])])[:-1] + rank * self.per
        self._attached = {}
        # page-locked landing buffer of the one download per step, and the positions of records / status
        # records in it (a pure function of the counts); the order of the ids is learnt from the first gather
        self.host = np.zeros(world * self.per, dtype=RECORD)
        capi.pin(ctx, self.host)
        rec_pos, status_pos = [], []
        for r, seg in enumerate(self.counts):
            pos = r * self.per
            for cnt in seg:
                rec_pos.append(np.arange(pos, pos + cnt))
                status_pos.append(pos + cnt)
                pos += cnt + 1
        self.rec_pos = np.concatenate(rec_pos) if rec_pos else np.zeros(0, dtype=np.int64)
        self.status_pos = np.asarray(status_pos, dtype=np.int64)
        self.perm = None

    @property
    def d2h_bytes(self) -> int:
        return self.world * self.per * RECORD.itemsize

    def attach(self, batch, ids, k: int = 0):
        """Point the finalize phase of sub-batch k at its slots of this rank's slice."""
        if self._attached.get(k) is not batch:
            assert len(ids) == self.counts[self.rank][k]
            batch.set_record_ids(ids)
            batch.set_record_sink(self.buf, int(self.first[k]))
            self._attached[k] = batch

    def launch_gather(self):
        """Enqueue the in-place all-gather on the context stream, behind the kernels that fill the slice."""
        if self.comm is not None:
            self.comm.all_gather_records(self.buf, self.per)

    def gather(self):
        """All-gather on the context stream (in place), one download, assembly in global order."""
        self.launch_gather()
        return self.collect()

    def collect(self):
        """One download of the gathered buffer (waits for the context stream), assembly in global order."""
        from . import capi
        capi.download_into(self.ctx, self.buf, self.host)
        ok = bool((self.host["found"][self.status_pos] == 1).all())
        if self.perm is None:               # global id -> position in the gathered buffer, fixed for this exchange
            ids = self.host["submap"][self.rec_pos]
            perm = np.empty(self.n_items, dtype=np.int64)
            perm[ids] = self.rec_pos
            self.perm = perm
        return self.host[self.perm], ok

    def finish(self, batches):
        """gather(); if any rank's run has to be repeated exactly (its status record says so), every rank
        sees that in the gathered buffer, the ranks settle their batches (exact path, records rewritten
        in place) and the gather is repeated -- a collective decision taken from exchanged data."""
        return self.finish_launched(batches, launched=False)

    def finish_launched(self, batches, launched=True):
        """finish() for a caller that has already enqueued the all-gather (launch_gather) behind its kernels."""
        out, ok = self.collect() if launched else self.gather()
        if not ok:
            for b in batches:
                b.settle()
            out, ok = self.gather()
            assert ok
        return out

    def step(self, batch, scans, pair_scan, pyramids, ids, thr):
        """One sharded loop-detection step: upload, ONE kernel launch, in-place all-gather, download."""
        self.attach(batch, ids)
        batch.upload_pairs(scans, pair_scan, pyramids, thr)
        batch.run()
        return self.finish([batch])

    def close(self):
        from . import capi
        if self.buf:
            capi.unpin(self.ctx, self.host)
            capi.device_free(self.ctx, self.buf)
            self.buf = 0


def rank_grid(world: int, n_submaps: int, submaps_per_group: int = 250, want_pm: int | None = None):
    """(Ps, Pm): scan groups x submap groups with Ps * Pm == world.  (scan, submap) pairs are independent,
    so a batch of scans against many submaps can be split along both axes; Pm is the largest divisor of
    `world` that still leaves about `submaps_per_group` submaps (a good device sub-batch) per group."""
    want = max(1, n_submaps // max(submaps_per_group, 1)) if want_pm is None else max(1, int(want_pm))
    pm = max(d for d in range(1, world + 1) if world % d == 0 and d <= want)
    return world // pm, pm


def grid_owned(n_scans: int, n_submaps: int, rank: int, ps_groups: int, pm_groups: int):
    """Scans and submaps of `rank` on a ps_groups x pm_groups rank grid (rank = ps * pm_groups + pm):
    scans k with k % ps_groups == ps, submaps g with g % pm_groups == pm.  Since pm_groups divides the
    world size, the submaps a rank owns under plain round-robin (g % world == rank) are among them."""
    pm, ps = rank % pm_groups, rank // pm_groups
    scans = np.arange(n_scans, dtype=np.int64)
    submaps = np.arange(n_submaps, dtype=np.int64)
    return scans[scans % ps_groups == ps], submaps[submaps % pm_groups == pm]


def best_candidate(records: np.ndarray):
    """Index of the found record with the highest score (ties: lowest submap index), or -1."""
    ok = np.flatnonzero(records["found"] != 0)
    if len(ok) == 0:
        return -1
    return int(ok[np.argmax(records["score"][ok])])
