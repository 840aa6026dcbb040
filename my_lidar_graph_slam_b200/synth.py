"""Deterministic synthetic 2D worlds and laser scans.

Axis-aligned rectangular rooms/corridors with analytic ray casting and Gaussian range
noise.  The same scans feed the oracle and the CUDA path, so everything here is plain
numpy with explicit seeds (world layout 1, range noise 2, pose perturbations 3).
This is benchmark / test tooling, not part of the device hot path.
"""
from __future__ import annotations

import numpy as np


class World:
    """A set of axis-aligned wall segments inside an outer rectangle."""

    def __init__(self, size_x: float = 40.0, size_y: float = 40.0, n_boxes: int = 12,
                 seed: int = 1):
        rng = np.random.default_rng(seed)
        self.size_x, self.size_y = float(size_x), float(size_y)
        hx, hy = size_x / 2.0, size_y / 2.0
        segs = [(-hx, -hy, hx, -hy), (hx, -hy, hx, hy), (hx, hy, -hx, hy), (-hx, hy, -hx, -hy)]
        self.boxes = []
        for _ in range(n_boxes):
            w, h = rng.uniform(1.0, 0.2 * size_x), rng.uniform(1.0, 0.2 * size_y)
            cx = rng.uniform(-hx + w / 2 + 1.0, hx - w / 2 - 1.0)
            cy = rng.uniform(-hy + h / 2 + 1.0, hy - h / 2 - 1.0)
            x0, x1, y0, y1 = cx - w / 2, cx + w / 2, cy - h / 2, cy + h / 2
            self.boxes.append((x0, y0, x1, y1))
            segs += [(x0, y0, x1, y0), (x1, y0, x1, y1), (x1, y1, x0, y1), (x0, y1, x0, y0)]
        s = np.asarray(segs, dtype=np.float64)
        self.horizontal = s[s[:, 1] == s[:, 3]]   # y = const
        self.vertical = s[s[:, 0] == s[:, 2]]     # x = const

    def is_free(self, x: float, y: float, margin: float = 0.4) -> bool:
        if abs(x) > self.size_x / 2 - margin or abs(y) > self.size_y / 2 - margin:
            return False
        for (x0, y0, x1, y1) in self.boxes:
            if x0 - margin <= x <= x1 + margin and y0 - margin <= y <= y1 + margin:
                return False
        return True

    def _lines(self):
        """Group collinear wall pieces: {coordinate: sorted interval end points} per axis."""
        if getattr(self, "_line_cache", None) is None:
            out = []
            for segs, fixed, lo, hi in ((self.horizontal, 1, 0, 2), (self.vertical, 0, 1, 3)):
                groups = {}
                for sg in segs:
                    groups.setdefault(float(sg[fixed]), []).append((min(sg[lo], sg[hi]), max(sg[lo], sg[hi])))
                lines = []
                for coord, iv in sorted(groups.items()):
                    iv.sort()
                    merged = [list(iv[0])]
                    for a, b in iv[1:]:
                        if a <= merged[-1][1]:
                            merged[-1][1] = max(merged[-1][1], b)
                        else:
                            merged.append([a, b])
                    lines.append((coord, np.asarray(merged).reshape(-1)))
                out.append(lines)
            self._line_cache = out
        return self._line_cache

    def cast(self, x: float, y: float, angles: np.ndarray) -> np.ndarray:
        """Distance to the first wall along each absolute angle."""
        c, s = np.cos(angles), np.sin(angles)
        best = np.full(angles.shape, 1e9)
        hor, ver = self._lines()
        with np.errstate(divide="ignore", invalid="ignore"):
            for coord, ends in hor:           # wall pieces on the line y = coord
                t = (coord - y) / s
                inside = np.searchsorted(ends, x + t * c, side="right") % 2 == 1
                best = np.where(inside & (t > 1e-9) & (t < best), t, best)
            for coord, ends in ver:           # wall pieces on the line x = coord
                t = (coord - x) / c
                inside = np.searchsorted(ends, y + t * s, side="right") % 2 == 1
                best = np.where(inside & (t > 1e-9) & (t < best), t, best)
        return best


def beam_angles(n_beams: int = 1081, fov_deg: float = 270.0) -> np.ndarray:
    fov = np.deg2rad(fov_deg)
    return np.ascontiguousarray(-fov / 2.0 + fov * np.arange(n_beams) / (n_beams - 1))


def make_scan(world: World, pose, angles: np.ndarray, noise_rng: np.random.Generator | None,
              sigma: float = 0.01, max_range: float = 30.0) -> np.ndarray:
    """Ranges of one scan taken with the *sensor* at `pose` = (x, y, theta)."""
    r = world.cast(pose[0], pose[1], pose[2] + angles)
    if noise_rng is not None:
        r = r + noise_rng.normal(0.0, sigma, size=r.shape)
    return np.ascontiguousarray(np.clip(r, 0.02, max_range))


def trajectory(world: World, n: int, step: float = 0.25, seed: int = 1,
               start=None) -> np.ndarray:
    """A smooth random walk of `n` poses that stays in free space."""
    rng = np.random.default_rng(seed + 1000)
    if start is None:
        while True:
            x, y = rng.uniform(-world.size_x / 2, world.size_x / 2), rng.uniform(
                -world.size_y / 2, world.size_y / 2)
            if world.is_free(x, y, 0.8):
                break
        th = rng.uniform(-np.pi, np.pi)
    else:
        x, y, th = start
    out = np.empty((n, 3))
    for k in range(n):
        out[k] = (x, y, th)
        for _ in range(64):
            nth = th + rng.normal(0.0, 0.15)
            nx, ny = x + step * np.cos(nth), y + step * np.sin(nth)
            if world.is_free(nx, ny, 0.6):
                x, y, th = nx, ny, nth
                break
            th += rng.uniform(0.5, 1.5)
    return out


def rasterize_map(scan_poses, angles, ranges_list, res: float = 0.05, patch: int = 64,
                  rmin: float = 0.01, rmax: float = 20.0, p_hit: float = 0.6,
                  p_miss: float = 0.45):
    """Numpy stand-in for scan integration, used ONLY to synthesise benchmark grids without
    touching the oracle: per-cell hit/miss counts folded through the odds product, clamped to
    [1e-3, 0.999], 0.0 = unknown, geometry aligned to `patch` cells.  Order-insensitive, so it
    is not bit-identical to the reference integration -- it only reproduces its value
    distribution (free ~1e-3..0.45, walls ~0.6..0.999).  Returns (dense[ny][nx], min_x, min_y).
    """
    pts = []
    for pose, r in zip(scan_poses, ranges_list):
        ok = (r > rmin) & (r < rmax)
        pts.append(np.stack([pose[0] + r[ok] * np.cos(pose[2] + angles[ok]),
                             pose[1] + r[ok] * np.sin(pose[2] + angles[ok])], 1))
        pts.append(np.asarray(pose[:2])[None, :])
    allp = np.concatenate(pts)
    lo = np.floor(allp.min(0) / (res * patch)) * (res * patch)
    hi = np.ceil((allp.max(0) + 1e-9) / (res * patch)) * (res * patch)
    nx, ny = int(round((hi[0] - lo[0]) / res)), int(round((hi[1] - lo[1]) / res))
    hits = np.zeros((ny, nx), dtype=np.int32)
    miss = np.zeros((ny, nx), dtype=np.int32)
    for pose, r in zip(scan_poses, ranges_list):
        ok = (r > rmin) & (r < rmax)
        rr, aa = r[ok], pose[2] + angles[ok]
        c, s = np.cos(aa), np.sin(aa)
        ex = np.floor((pose[0] + rr * c - lo[0]) / res).astype(np.int64)
        ey = np.floor((pose[1] + rr * s - lo[1]) / res).astype(np.int64)
        np.add.at(hits, (ey, ex), 1)
        nstep = int(np.ceil(rr.max() / (0.5 * res)))
        t = (np.arange(nstep) * 0.5 * res)[None, :]
        valid = t < (rr[:, None] - res)
        mx = np.floor((pose[0] + t * c[:, None] - lo[0]) / res).astype(np.int64)
        my = np.floor((pose[1] + t * s[:, None] - lo[1]) / res).astype(np.int64)
        lin = (np.arange(len(rr))[:, None] * (nx * ny) + my * nx + mx)[valid]
        lin = np.unique(lin) % (nx * ny)          # one miss per (beam, cell)
        np.add.at(miss.reshape(-1), lin, 1)
    odds = (p_hit / (1 - p_hit)) ** hits * (p_miss / (1 - p_miss)) ** miss
    dense = np.clip(odds / (1 + odds), 1e-3, 1 - 1e-3)
    dense[(hits + miss) == 0] = 0.0
    return np.ascontiguousarray(dense), float(lo[0]), float(lo[1])


class RoomsWorld(World):
    """Office-like world: a lattice of `room` x `room` metre rooms whose walls each have one
    door gap.  Most beams end within one room (ranges below ~room * 1.4) while a few pass
    through doors, which is the range distribution config C2 assumes (ScanRangeMax 5.7 m keeps
    nearly every beam) on a map that still spans hundreds of cells."""

    def __init__(self, size: float = 40.0, room: float = 5.0, door: float = 1.2, seed: int = 1):
        rng = np.random.default_rng(seed)
        self.size_x = self.size_y = float(size)
        self.boxes = []
        h = size / 2.0
        n = int(round(size / room))
        segs = [(-h, -h, h, -h), (h, -h, h, h), (h, h, -h, h), (-h, h, -h, -h)]
        for i in range(1, n):          # interior wall lines
            c = -h + i * room
            for j in range(n):         # one wall piece per room side, with a door gap
                a0 = -h + j * room
                g = a0 + rng.uniform(0.4, room - door - 0.4)
                segs += [(c, a0, c, g), (c, g + door, c, a0 + room)]          # vertical wall x = c
                g = a0 + rng.uniform(0.4, room - door - 0.4)
                segs += [(a0, c, g, c), (g + door, c, a0 + room, c)]          # horizontal wall y = c
        s = np.asarray(segs, dtype=np.float64)
        self.horizontal = s[s[:, 1] == s[:, 3]]
        self.vertical = s[s[:, 0] == s[:, 2]]
        self.room, self.n_rooms = room, n

    def is_free(self, x: float, y: float, margin: float = 0.4) -> bool:
        h = self.size_x / 2.0
        if abs(x) > h - margin or abs(y) > h - margin:
            return False
        fx = (x + h) % self.room
        fy = (y + h) % self.room
        return (margin <= fx <= self.room - margin) and (margin <= fy <= self.room - margin)
