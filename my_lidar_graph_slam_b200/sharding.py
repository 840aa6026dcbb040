"""Multi-GPU partitioning of loop-detection work.

(scan, submap) branch-and-bound queries are independent, so submap i (with its device-resident
pyramid) lives on rank i mod world and every rank searches only its own pairs.  The one exchange
step is an all-gather of the fixed-size result records (found, ix, iy, itheta, score), after
which every rank holds the full result list in global submap order and can pick the best
candidate.  torch.distributed is plumbing only: NCCL over NVLink on the GPU box, gloo in the CPU
tests.
"""
from __future__ import annotations

import numpy as np

RECORD = np.dtype([("found", np.int32), ("ix", np.int32), ("iy", np.int32), ("it", np.int32),
                   ("score", np.float64), ("submap", np.int64)])   # 32 bytes


def owned(n_items: int, rank: int, world: int) -> np.ndarray:
    """Global indices owned by `rank` under round-robin placement."""
    return np.arange(rank, n_items, world, dtype=np.int64)


def pack(results, global_ids) -> np.ndarray:
    rec = np.zeros(len(global_ids), dtype=RECORD)
    for k, (r, g) in enumerate(zip(results, global_ids)):
        rec[k] = (r.found, r.ix, r.iy, r.it, r.score, g)
    return rec


def pack_array(results: np.ndarray, global_ids) -> np.ndarray:
    """pack() for a structured result array (capi.BbBatch.results_array), vectorised."""
    rec = np.zeros(len(global_ids), dtype=RECORD)
    for f in ("found", "ix", "iy", "it", "score"):
        rec[f] = results[f]
    rec["submap"] = global_ids
    return rec


def all_gather_records(local: np.ndarray, n_items: int, rank: int, world: int, device=None):
    """All-gather per-rank record arrays -> one array of n_items records in global order."""
    if world == 1:
        out = np.zeros(n_items, dtype=RECORD)
        out[local["submap"]] = local
        return out
    import torch
    import torch.distributed as dist
    per = (n_items + world - 1) // world          # ranks own per or per-1 items: pad to per
    buf = np.zeros(per, dtype=RECORD)
    buf["submap"] = -1
    buf[:len(local)] = local
    send = torch.from_numpy(buf.view(np.uint8).copy())
    if device is not None:
        send = send.to(device)
    recv = torch.empty(world * send.numel(), dtype=torch.uint8, device=send.device)
    dist.all_gather_into_tensor(recv, send)
    allrec = recv.cpu().numpy().view(RECORD)
    allrec = allrec[allrec["submap"] >= 0]
    out = np.zeros(n_items, dtype=RECORD)
    out[allrec["submap"]] = allrec
    return out


def all_gather_variable(local: np.ndarray, n_items: int, world: int, device=None):
    """All-gather record arrays whose per-rank counts are arbitrary (queries follow the band their
    sensor cell falls in): one all-reduce(MAX) of the counts, then one padded all-gather."""
    if world == 1:
        out = np.zeros(n_items, dtype=RECORD)
        out[local["submap"]] = local
        return out
    import torch
    import torch.distributed as dist
    cnt = torch.tensor([len(local)], dtype=torch.int64, device=device if device is not None else "cpu")
    dist.all_reduce(cnt, op=dist.ReduceOp.MAX)
    per = max(int(cnt.item()), 1)
    buf = np.zeros(per, dtype=RECORD)
    buf["submap"] = -1
    buf[:len(local)] = local
    send = torch.from_numpy(buf.view(np.uint8).copy())
    if device is not None:
        send = send.to(device)
    recv = torch.empty(world * send.numel(), dtype=torch.uint8, device=send.device)
    dist.all_gather_into_tensor(recv, send)
    allrec = recv.cpu().numpy().view(RECORD)
    allrec = allrec[allrec["submap"] >= 0]
    out = np.zeros(n_items, dtype=RECORD)
    out[allrec["submap"]] = allrec
    return out


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
