"""Large-map loop closure split into row bands, one band per GPU.

One big occupancy grid (e.g. 8000 x 8000 cells) is cut into `world` bands of rows.  Rank g keeps the
rows of band g plus a margin on both sides, declared to the library as a WINDOW of the whole map
(lgs_grid_set_window): world -> cell conversion stays floor((p - min) / res) of the whole map, so
every projected index -- and therefore every score and winner -- is bit-identical to matching
against the whole map on one device.  The multi-resolution pyramid
(PrecomputeGridMaps, mapping/grid_map_builder.cpp:471-495) is built per band from the band's own
rows: out(x, y) = max in[x .. x + w) x [y .. y + w) only looks UP and RIGHT, so a band needs
2^H - 1 extra rows above the highest row a match can read (that is the one-directional halo of
SURVEY 8(e)); here the host uploads band + margin directly, so no device halo exchange is needed.
The reference's clamped last window (xs = min(x, n - w)) only differs from the whole
map inside the top 2^H - 1 rows of a band that does not end at the map's edge, and the margin keeps
every read below them.

A loop query (scan, initial pose) goes to the rank whose band contains the row of its sensor cell;
per-rank results are all-gathered as fixed-size records (sharding.all_gather_records).

Rows a branch-and-bound match can read (ScanMatcherBranchBound::OptimizePose,
mapping/scan_matcher_branch_bound.cpp:47-163): sensor row + node offset in [-winY, winY + 2^H)
+ beam reach (ranges above the usable maximum are skipped, score_function_pixel_accurate.cpp:36-44).
"""
from __future__ import annotations

import math

import numpy as np


def band_rows(ny: int, world: int, rank: int):
    """Rows [r0, r1) owned by `rank`: equal bands, the remainder spread over the first ranks."""
    base, extra = divmod(ny, world)
    r0 = rank * base + min(rank, extra)
    return r0, r0 + base + (1 if rank < extra else 0)


def margins(res: float, reach_m: float, range_y_m: float, height_max: int):
    """(below, above) margins in rows so that no match of a scan whose sensor cell lies in the band
    reads outside the window and no read touches a clamped pyramid window."""
    reach = int(math.ceil(reach_m / res)) + 2           # beam end + rounding of the pose / cell
    win = int(math.ceil(0.5 * range_y_m / res))          # winY (scan_matcher_branch_bound.cpp:68-73)
    node = 1 << height_max
    below = reach + win + 2
    above = reach + win + node + (node - 1) + 2          # node span, then the pyramid's look-ahead
    return below, above


def window_rows(ny: int, world: int, rank: int, below: int, above: int):
    r0, r1 = band_rows(ny, world, rank)
    return max(r0 - below, 0), min(r1 + above, ny)


def owner_of_rows(sensor_rows, ny: int, world: int) -> np.ndarray:
    """Rank owning each sensor row (rows outside the map go to the nearest band)."""
    rows = np.clip(np.asarray(sensor_rows, dtype=np.int64), 0, ny - 1)
    edges = np.array([band_rows(ny, world, g)[1] for g in range(world)], dtype=np.int64)
    return np.searchsorted(edges, rows, side="right").astype(np.int64)


def sensor_rows(sensor_y, min_y: float, res: float) -> np.ndarray:
    """floor((y - min_y) / res) like GridMap::WorldCoordinateToGridCellIndex (grid_map.hpp:779-790)."""
    return np.floor((np.asarray(sensor_y, dtype=np.float64) - min_y) / res).astype(np.int64)


class BandedMap:
    """Band `rank` of `world` of a big map, resident on one device with its pyramid."""

    def __init__(self, ctx, rows_provider, nx, ny, min_x, min_y, res, rank, world, height_max,
                 reach_m, range_y_m, apron=1):
        from . import capi
        self.rank, self.world, self.ny_total = rank, world, ny
        below, above = margins(res, reach_m, range_y_m, height_max)
        self.r0, self.r1 = band_rows(ny, world, rank)
        self.w0, self.w1 = window_rows(ny, world, rank, below, above)
        rows = np.ascontiguousarray(rows_provider(self.w0, self.w1), dtype=np.float64)
        assert rows.shape == (self.w1 - self.w0, nx), rows.shape
        self.grid = capi.Grid(ctx, nx, self.w1 - self.w0, min_x, min_y, res, apron=apron)
        self.grid.set_window(0, self.w0)
        self.grid.upload(rows)
        self.height_max = height_max
        self.pyramid = None
        self.ctx = ctx

    def build_pyramid(self):
        from . import capi
        self.pyramid = capi.Pyramid(self.ctx, self.grid, self.height_max)
        return self.pyramid

    @property
    def cells(self):
        return self.grid.nx * self.grid.ny

    def close(self):
        if self.pyramid is not None:
            self.pyramid.close()
        self.grid.close()
