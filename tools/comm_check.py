"""Run under torchrun: loop detection sharded over one process per GPU, records all-gathered in place through
lgs_comm (NCCL on the context stream); every rank checks the gathered records against its own single-GPU
run of ALL pairs."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
from my_lidar_graph_slam_b200 import capi, sharding, synth  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
ctx = capi.Context(local)
comm = sharding.make_comm(ctx, rank, world)

N_SUB = 11
world_map = synth.RoomsWorld(40.0, 5.0, seed=6)
angles = synth.beam_angles(1081, 270.0)
anchor = synth.trajectory(world_map, 1, seed=31)[0]
pyr = []
for g in range(N_SUB):
    rng = np.random.default_rng(500 + g)
    traj = synth.trajectory(world_map, 8, step=0.3, seed=g, start=tuple(anchor) if g % 2 == 0 else None) \
        if g % 2 == 0 else synth.trajectory(world_map, 8, step=0.3, seed=g)
    scans = [synth.make_scan(world_map, p, angles, rng) for p in traj]
    grid, _ = bench.build_map_on_gpu(ctx, traj, angles, scans, apron=1)
    pyr.append(capi.Pyramid(ctx, grid, 6))
qscan = synth.make_scan(world_map, anchor, angles, np.random.default_rng(9))
init = anchor + np.array([0.3, -0.2, 0.06])
scans = capi.Scans([angles], [qscan], [init], range_min=0.02, range_max=30.0)
# all pairs on this rank alone
full = capi.BbBatch(ctx, **bench.BB)
full.upload_pairs(scans, np.zeros(N_SUB, dtype=np.int32), pyr, 0.55)
full.run()
want = full.records()
# this rank's share, exchanged on the device
mine = sharding.owned(N_SUB, rank, world)
ex = sharding.RecordExchange(ctx, comm, N_SUB, rank, world)
part = capi.BbBatch(ctx, **bench.BB)
for rep in range(3):
    got = ex.step(part, scans, np.zeros(len(mine), dtype=np.int32), [pyr[int(g)] for g in mine], mine, 0.55)
    assert got.tobytes() == want.tobytes(), (rank, rep)
dist.barrier()
if rank == 0:
    print(f"COMM_CHECK_OK world={world} found={int((got['found'] != 0).sum())}")
dist.destroy_process_group()
