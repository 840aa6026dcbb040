"""CPU: the N > 1 loop-detection path (round-robin submap placement + all-gather of result records)
on world_size 2 with the gloo backend."""
import os
import socket
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_items, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from my_lidar_graph_slam_b200 import sharding
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    ids = sharding.owned(n_items, rank, world)

    class R:   # stand-in for lgs_match_result
        def __init__(self, g):
            self.found, self.ix, self.iy, self.it, self.score = int(g % 3 == 0), g, -g, 2 * g, 100.0 + (g * 7) % 11
    local = sharding.pack([R(int(g)) for g in ids], ids)
    full = sharding.all_gather_records(local, n_items, rank, world)
    q.put((rank, full.tobytes(), sharding.best_candidate(full)))
    dist.destroy_process_group()


def test_round_robin_ownership_partitions_everything():
    from my_lidar_graph_slam_b200 import sharding
    for n in (0, 1, 7, 500):
        for world in (1, 2, 4, 8):
            parts = [sharding.owned(n, r, world) for r in range(world)]
            assert sorted(np.concatenate(parts).tolist()) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_all_gather_records_world2_gloo():
    import torch.multiprocessing as mp
    from my_lidar_graph_slam_b200 import sharding
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port, n_items = _free_port(), 7           # odd count: ranks own 4 and 3 records
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_items, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full0 = np.frombuffer(got[0][1], dtype=sharding.RECORD)
    assert got[0][1] == got[1][1] and got[0][2] == got[1][2]      # every rank holds the same list
    assert full0["submap"].tolist() == list(range(n_items))
    assert full0["ix"].tolist() == list(range(n_items)) and full0["found"].tolist() == [1, 0, 0, 1, 0, 0, 1]
    best = got[0][2]
    assert full0["found"][best] == 1 and full0["score"][best] == max(full0["score"][[0, 3, 6]])
