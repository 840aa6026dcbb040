"""Seeded synthetic scenes shared by the parity tests (maps are built by the oracle)."""
from __future__ import annotations

import functools

import numpy as np

from my_lidar_graph_slam_b200 import synth
from oracle import backend

R = backend()


@functools.lru_cache(maxsize=8)
def room_scene(seed: int = 1, n_map_scans: int = 10, n_beams: int = 1081, fov: float = 270.0,
               size: float = 24.0, n_boxes: int = 10):
    """World + reference-built latest map from the first scans of a trajectory."""
    world = synth.World(size, size, n_boxes, seed=seed)
    angles = synth.beam_angles(n_beams, fov)
    traj = synth.trajectory(world, n_map_scans + 8, step=0.2, seed=seed)
    noise = np.random.default_rng(seed + 1)
    builder = R.RefBuilder()
    for p in traj[:n_map_scans]:
        builder.append_scan(p, angles, synth.make_scan(world, p, angles, noise))
    return world, angles, traj, builder
