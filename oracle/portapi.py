"""ctypes binding to oracle/_build/liblgs_oracle.so -- TEST INFRASTRUCTURE ONLY.

The plain-C restatement of the reference hot path (oracle/lgs_oracle.c) behind the same Python
surface as oracle/refapi.py, so every parity test can run against either backend:
`from oracle import backend; R = backend()` picks the unmodified reference objects when
oracle/_ref/liblgs_ref.so exists and this port otherwise.  The builder logic that is plain
bookkeeping in the reference (GridMapBuilder::AppendScan's local-map / latest-map sequencing,
grid_map_builder.cpp:48-60, :98-207) is restated here in Python on top of the C primitives.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "liblgs_oracle.so")

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)
TINY = float(np.finfo(np.float64).tiny)


class Geom(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("min_x", C.c_double), ("min_y", C.c_double),
                ("res", C.c_double), ("patch", C.c_int)]


class PortMatch(C.Structure):
    _fields_ = [("found", C.c_int), ("ix", C.c_int), ("iy", C.c_int), ("it", C.c_int),
                ("winX", C.c_int), ("winY", C.c_int), ("winT", C.c_int), ("pad", C.c_int),
                ("stepX", C.c_double), ("stepY", C.c_double), ("stepT", C.c_double),
                ("score", C.c_double), ("sensorPose", C.c_double * 3),
                ("bestSensorPose", C.c_double * 3), ("n_scored", C.c_longlong)]


def available() -> bool:
    return os.path.exists(LIB_PATH)


def build():
    subprocess.run(["make", "port"], cwd=_HERE, check=True, stdout=subprocess.PIPE,
                   stderr=subprocess.STDOUT)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not available():
            build()
        L = C.CDLL(LIB_PATH)
        G = C.POINTER(Geom)
        L.orc_bayes_update.restype = C.c_double
        L.orc_bayes_update.argtypes = [C.c_double, C.c_double]
        L.orc_bresenham.restype = C.c_int
        L.orc_bresenham.argtypes = [C.c_int] * 4 + [c_ip, C.c_int]
        L.orc_sliding_window_max.argtypes = [c_dp, C.c_int, C.c_int, c_dp]
        L.orc_precompute.argtypes = [c_dp, C.c_int, C.c_int, C.c_int, c_dp]
        L.orc_pyramid.argtypes = [c_dp, C.c_int, C.c_int, C.c_int, c_dp]
        L.orc_compound.argtypes = [c_dp, c_dp, c_dp]
        L.orc_hit_points.restype = C.c_int
        L.orc_hit_points.argtypes = [c_dp, c_dp, C.c_int, c_dp, c_dp] + [C.c_double] * 4 + [c_dp] * 3
        L.orc_geometry_resize.argtypes = [G] + [C.c_double] * 4 + [G, c_ip, c_ip]
        L.orc_geometry_expand.restype = C.c_int
        L.orc_geometry_expand.argtypes = [G] + [C.c_double] * 5 + [G, c_ip, c_ip]
        L.orc_integrate_hits.restype = C.c_int
        L.orc_integrate_hits.argtypes = [c_dp, G, c_dp, C.c_int, c_dp, C.c_double, C.c_double]
        L.orc_rtcsm_match.restype = C.c_int
        L.orc_rtcsm_match.argtypes = [c_dp, c_dp, G, C.c_int] + [C.c_double] * 4 + \
            [c_dp, c_dp, C.c_int, c_dp, c_dp, C.c_double, C.POINTER(PortMatch)]
        L.orc_pixel_accurate_score.restype = C.c_double
        L.orc_pixel_accurate_score.argtypes = [c_dp, G, C.c_double, C.c_double, c_dp, C.c_int, c_dp,
                                               c_dp, C.c_double, C.c_double]
        L.orc_cost_greedy_endpoint.restype = C.c_double
        L.orc_cost_greedy_endpoint.argtypes = [c_dp, G, c_dp, c_dp, C.c_int, c_dp, c_dp, C.c_double, C.c_double]
        L.orc_cost_tail.argtypes = [c_dp, G, c_dp, c_dp, C.c_int, c_dp, c_dp, C.c_double, C.c_double,
                                    c_dp, c_dp]
        L.orc_bb_match.restype = C.c_int
        L.orc_bb_match.argtypes = [c_dp, G, C.c_int] + [C.c_double] * 6 + \
            [c_dp, c_dp, C.c_int, c_dp, c_dp, C.c_double, C.c_double, C.c_double, C.POINTER(PortMatch)]
        _lib = L
    return _lib


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(c_dp)


def _arr3(p):
    return (C.c_double * 3)(*[float(v) for v in p])


class PortPre:
    def __init__(self, dense):
        self._dense = dense

    def dense(self):
        return self._dense.copy()


class PortMap:
    """Dense stand-in for GridMap<BinaryBayesGridCell<double>>."""

    def __init__(self, dense, min_x, min_y, res=0.05, patch=64):
        self.d = np.ascontiguousarray(dense, dtype=np.float64)
        self.geo = Geom(self.d.shape[1], self.d.shape[0], float(min_x), float(min_y), float(res), int(patch))

    @classmethod
    def from_dense(cls, dense, min_x, min_y, res=0.05, patch=64):
        dense = np.asarray(dense, dtype=np.float64)
        assert dense.shape[0] % patch == 0 and dense.shape[1] % patch == 0
        return cls(dense.copy(), min_x, min_y, res, patch)

    def geometry(self):
        return self.geo.nx, self.geo.ny, self.geo.min_x, self.geo.min_y, self.geo.res

    def dense(self):
        return self.d.copy()

    def precompute(self, win):
        out = np.empty_like(self.d)
        if self.d.size:
            lib().orc_precompute(self.d.ctypes.data_as(c_dp), self.geo.nx, self.geo.ny, int(win),
                                 out.ctypes.data_as(c_dp))
        return PortPre(out)

    def pyramid(self, height_max):
        out = np.empty((height_max + 1,) + self.d.shape)
        if self.d.size:
            lib().orc_pyramid(self.d.ctypes.data_as(c_dp), self.geo.nx, self.geo.ny, int(height_max),
                              out.ctypes.data_as(c_dp))
        return [PortPre(out[h]) for h in range(height_max + 1)]

    def _apply_geometry(self, new: Geom, sx: int, sy: int):
        nd = np.zeros((new.ny, new.nx))
        ox0, oy0 = max(0, sx), max(0, sy)                      # overlap in old coordinates
        ox1, oy1 = min(self.geo.nx, new.nx + sx), min(self.geo.ny, new.ny + sy)
        if ox1 > ox0 and oy1 > oy0:
            nd[oy0 - sy:oy1 - sy, ox0 - sx:ox1 - sx] = self.d[oy0:oy1, ox0:ox1]
        self.d, self.geo = nd, new


def empty_map(center_x, center_y, res=0.05, patch=64) -> PortMap:
    return PortMap(np.zeros((0, 0)), center_x, center_y, res, patch)


def map_resize(m: PortMap, min_x, min_y, max_x, max_y):
    new, sx, sy = Geom(), C.c_int(), C.c_int()
    lib().orc_geometry_resize(C.byref(m.geo), min_x, min_y, max_x, max_y, C.byref(new), C.byref(sx), C.byref(sy))
    m._apply_geometry(new, sx.value, sy.value)


def map_expand(m: PortMap, min_x, min_y, max_x, max_y, step=5.0):
    new, sx, sy = Geom(), C.c_int(), C.c_int()
    if lib().orc_geometry_expand(C.byref(m.geo), min_x, min_y, max_x, max_y, step, C.byref(new),
                                 C.byref(sx), C.byref(sy)):
        m._apply_geometry(new, sx.value, sy.value)


def map_reset(m: PortMap):
    m.d[:] = 0.0


def map_integrate_hits(m: PortMap, sensor_xy, hit_xy, p_hit=0.6, p_miss=0.45) -> int:
    s, sp = _d(np.asarray(sensor_xy, dtype=np.float64)[:2])
    h, hp = _d(hit_xy)
    n = lib().orc_integrate_hits(m.d.ctypes.data_as(c_dp), C.byref(m.geo), sp, len(h), hp, p_hit, p_miss)
    if n < 0:
        raise RuntimeError("a touched cell lies outside the oracle map")
    return n


def hit_points(robot_pose, angles, ranges, *, rel=(0.0, 0.0, 0.0), scan_min_range=0.02,
               scan_max_range=30.0, usable_min=0.01, usable_max=20.0):
    a, ap = _d(angles)
    r, rp = _d(ranges)
    sp = (C.c_double * 3)()
    bbox = (C.c_double * 4)()
    out = np.empty((len(a), 2))
    k = lib().orc_hit_points(_arr3(robot_pose), _arr3(rel), len(a), ap, rp, scan_min_range,
                             scan_max_range, usable_min, usable_max, sp, out.ctypes.data_as(c_dp), bbox)
    return np.array(sp), out[:k].copy(), np.array(bbox)


def bresenham(x0, y0, x1, y1):
    cap = abs(x1 - x0) + abs(y1 - y0) + 2
    buf = np.empty((cap, 2), dtype=np.int32)
    n = lib().orc_bresenham(x0, y0, x1, y1, buf.ctypes.data_as(c_ip), cap)
    return buf[:n].copy()


def bayes_update(v, p):
    return lib().orc_bayes_update(float(v), float(p))


def sliding_window_max(a, w):
    a, ap = _d(a)
    out = np.empty_like(a)
    lib().orc_sliding_window_max(ap, len(a), int(w), out.ctypes.data_as(c_dp))
    return out


def compound(a, b):
    o = (C.c_double * 3)()
    lib().orc_compound(_arr3(a), _arr3(b), o)
    return tuple(o)


class PortBuilder:
    """GridMapBuilder::AppendScan sequencing (grid_map_builder.cpp:48-60, :98-207) on PortMaps."""

    def __init__(self, res=0.05, patch=64, n_latest=10, travel_thr=20.0, rmin=0.01, rmax=20.0,
                 p_hit=0.6, p_miss=0.45, scan_min_range=0.02, scan_max_range=30.0,
                 rel_pose=(0.0, 0.0, 0.0)):
        self.res, self.patch, self.n_latest, self.travel_thr = res, patch, n_latest, travel_thr
        self.rmin, self.rmax, self.p_hit, self.p_miss = rmin, rmax, p_hit, p_miss
        self.scan_min_range, self.scan_max_range, self.rel = scan_min_range, scan_max_range, rel_pose
        self.nodes = []            # (pose, angles, ranges)
        self.local = []            # [PortMap, node_lo, node_hi]
        self.latest = PortMap(np.zeros((0, 0)), 0.0, 0.0, res, patch)
        self.travel_since = 0.0
        self.last_pose = None

    def _hits(self, k):
        pose, a, r = self.nodes[k]
        return hit_points(pose, a, r, rel=self.rel, scan_min_range=self.scan_min_range,
                          scan_max_range=self.scan_max_range, usable_min=self.rmin, usable_max=self.rmax)

    def append_node_only(self, pose, angles, ranges):
        self.nodes.append((np.asarray(pose, dtype=np.float64).copy(),
                           np.asarray(angles, dtype=np.float64), np.asarray(ranges, dtype=np.float64)))

    def update_grid_map(self) -> bool:                        # :98-193
        pose = self.nodes[-1][0]
        if self.local:
            # InverseCompound + Distance reduce to the Euclidean distance of the two poses
            # (pose.hpp:164-180: a rotation of (dx, dy) keeps its norm up to rounding; only the
            # 20 m threshold test consumes it)
            lp = self.last_pose
            s, c = np.sin(lp[2]), np.cos(lp[2])
            dx, dy = pose[0] - lp[0], pose[1] - lp[1]
            self.travel_since += float(np.sqrt((c * dx + s * dy) ** 2 + (-s * dx + c * dy) ** 2))
        self.last_pose = pose
        created = not self.local or self.travel_since >= self.travel_thr
        if created:
            self.local.append([PortMap(np.zeros((0, 0)), pose[0], pose[1], self.res, self.patch),
                               len(self.nodes) - 1, len(self.nodes) - 1])
            self.travel_since = 0.0
        m = self.local[-1][0]
        sp, hits, bbox = self._hits(len(self.nodes) - 1)
        map_expand(m, *bbox)
        map_integrate_hits(m, sp, hits, self.p_hit, self.p_miss)
        self.local[-1][2] = len(self.nodes) - 1
        return created

    def construct_into(self, m: PortMap, lo, hi):            # ConstructMapFromScans :227-332
        hp = [self._hits(k) for k in range(lo, hi + 1)]
        bl = [min(min(sp[0], b[0]) for sp, _, b in hp), min(min(sp[1], b[1]) for sp, _, b in hp)]
        tr = [max(max(sp[0], b[2]) for sp, _, b in hp), max(max(sp[1], b[3]) for sp, _, b in hp)]
        tr = [max(tr[0], TINY), max(tr[1], TINY)]            # :236-237 numeric_limits<double>::min()
        map_resize(m, bl[0], bl[1], tr[0], tr[1])
        map_reset(m)
        for sp, hits, _ in hp:
            map_integrate_hits(m, sp, hits, self.p_hit, self.p_miss)

    def update_latest_map(self):                              # :196-207
        hi = len(self.nodes) - 1
        self.construct_into(self.latest, max(0, hi - self.n_latest + 1), hi)

    def append_scan(self, pose, angles, ranges) -> bool:      # :48-60
        self.append_node_only(pose, angles, ranges)
        created = self.update_grid_map()
        self.update_latest_map()
        return created

    def construct_map(self, lo, hi) -> PortMap:
        m = PortMap(np.zeros((0, 0)), 0.0, 0.0, self.res, self.patch)
        self.construct_into(m, lo, hi)
        return m

    def num_local_maps(self):
        return len(self.local)

    def local_map(self, i) -> PortMap:
        m = self.local[i][0]
        return PortMap(m.d.copy(), m.geo.min_x, m.geo.min_y, m.geo.res, m.geo.patch)

    def local_map_nodes(self, i):
        return self.local[i][1], self.local[i][2]

    def latest_map(self) -> PortMap:
        m = self.latest
        return PortMap(m.d.copy(), m.geo.min_x, m.geo.min_y, m.geo.res, m.geo.patch)


def rtcsm_match(m: PortMap, angles, ranges, init_pose, *, low_res=5, range_x=1.0, range_y=1.0,
                range_theta=1.0471975512, scan_range_max=20.0, thr=None, pre: PortPre | None = None,
                rel=(0.0, 0.0, 0.0), **_ignored) -> PortMatch:
    a, ap = _d(angles)
    r, rp = _d(ranges)
    if pre is None:
        pre = m.precompute(low_res)
    out = PortMatch()
    lib().orc_rtcsm_match(m.d.ctypes.data_as(c_dp), pre._dense.ctypes.data_as(c_dp), C.byref(m.geo),
                          low_res, range_x, range_y, range_theta, scan_range_max, _arr3(init_pose),
                          _arr3(rel), len(a), ap, rp, TINY if thr is None else thr, C.byref(out))
    return out


DEFAULT_COST = (0.01, 20.0, 0.075, 0.1, 1.0, 0.05, 1.0)   # ctor order, as oracle/refapi.py


def cost_greedy_endpoint(m: PortMap, sensor_pose, angles, ranges, *, scan_min_range=0.02,
                         scan_max_range=30.0, cost=DEFAULT_COST) -> float:
    a, ap = _d(angles)
    r, rp = _d(ranges)
    return lib().orc_cost_greedy_endpoint(m.d.ctypes.data_as(c_dp), C.byref(m.geo), (C.c_double * 7)(*cost),
                                          _arr3(sensor_pose), len(a), ap, rp, scan_min_range, scan_max_range)


def host_tail(m: PortMap, best_sensor_pose, angles, ranges, *, scan_min_range=0.02, scan_max_range=30.0,
              cost=DEFAULT_COST):
    """(normalised cost, None, 3x3 covariance) of the matchers' host tail (refapi.host_tail's shape; the
    estimated pose in the middle is not restated here)."""
    a, ap = _d(angles)
    r, rp = _d(ranges)
    nc = C.c_double()
    cov = (C.c_double * 9)()
    lib().orc_cost_tail(m.d.ctypes.data_as(c_dp), C.byref(m.geo), (C.c_double * 7)(*cost),
                        _arr3(best_sensor_pose), len(a), ap, rp, scan_min_range, scan_max_range,
                        C.byref(nc), cov)
    return nc.value, None, np.array(cov).reshape(3, 3)


def rtcsm_score_table(m: PortMap, pre: PortPre, use_coarse, low_res, scan_range_max, sensor_pose,
                      angles, ranges, step_t, win_t, x_lo, nxw, y_lo, nyw, want_table=True):
    """Exhaustive scores through the port's projection (numpy gather, sequential beam order)."""
    a = np.asarray(angles, dtype=np.float64)
    r = np.asarray(ranges, dtype=np.float64)
    keep = r < scan_range_max
    nt = 2 * win_t + 1
    nx, ny, mx, my, res = m.geometry()
    src = pre._dense if use_coarse else m.d
    pad = np.zeros((ny + 2, nx + 2))
    pad[1:-1, 1:-1] = src
    idx = np.full((nt, len(a), 2), -(2 ** 31), dtype=np.int32)
    cnt = np.full(nt, int(keep.sum()), dtype=np.int32)
    table = np.zeros((nt, nyw, nxw)) if want_table else None
    oy, ox = np.meshgrid(np.arange(nyw) + y_lo, np.arange(nxw) + x_lo, indexing="ij")
    for t in range(-win_t, win_t + 1):
        pose = np.array([sensor_pose[0], sensor_pose[1], sensor_pose[2] + step_t * t])
        _, hits, _ = hit_points(pose, a[keep], r[keep], scan_min_range=-1.0, scan_max_range=np.inf,
                                usable_min=-1.0, usable_max=np.inf)
        cx = np.floor((hits[:, 0] - mx) / res).astype(np.int64)
        cy = np.floor((hits[:, 1] - my) / res).astype(np.int64)
        k = len(cx)
        idx[t + win_t, :k, 0], idx[t + win_t, :k, 1] = cx, cy
        if want_table:
            acc = np.zeros((nyw, nxw))
            for i in range(k):                                   # beam order, like ComputeScore
                acc += pad[np.clip(cy[i] + oy, -1, ny) + 1, np.clip(cx[i] + ox, -1, nx) + 1]
            table[t + win_t] = acc
    return table, idx, cnt


def bb_match(m: PortMap, angles, ranges, init_pose, *, height_max=6, range_x=2.0, range_y=2.0,
             range_theta=1.0, scan_range_max=20.0, score_range_min=0.01, score_range_max=20.0,
             thr=0.6, pyramid=None, rel=(0.0, 0.0, 0.0), scan_min_range=0.02, scan_max_range=30.0,
             **_ignored) -> PortMatch:
    a, ap = _d(angles)
    r, rp = _d(ranges)
    if pyramid is None:
        pyramid = m.pyramid(height_max)
    pyr = np.ascontiguousarray(np.stack([p._dense for p in pyramid]))
    out = PortMatch()
    lib().orc_bb_match(pyr.ctypes.data_as(c_dp), C.byref(m.geo), height_max, range_x, range_y,
                       range_theta, scan_range_max, score_range_min, score_range_max,
                       _arr3(init_pose), _arr3(rel), len(a), ap, rp, scan_min_range, scan_max_range,
                       thr, C.byref(out))
    return out


def gs_match(m: PortMap, angles, ranges, init_pose, *, range_x=2.0, range_y=2.0, range_theta=0.5,
             step_x=0.05, step_y=0.05, step_theta=0.005, score_range_min=0.01, score_range_max=20.0,
             thr=0.5, rel=(0.0, 0.0, 0.0), scan_min_range=0.02, scan_max_range=30.0,
             **_ignored) -> PortMatch:
    a, ap = _d(angles)
    r, rp = _d(ranges)
    out = PortMatch()
    L = lib()
    L.orc_gs_match.argtypes = [c_dp, C.POINTER(Geom)] + [C.c_double] * 8 + \
        [c_dp, c_dp, C.c_int, c_dp, c_dp, C.c_double, C.c_double, C.c_double, C.POINTER(PortMatch)]
    L.orc_gs_match(m.d.ctypes.data_as(c_dp), C.byref(m.geo), range_x, range_y, range_theta, step_x, step_y,
                   step_theta, score_range_min, score_range_max, _arr3(init_pose), _arr3(rel), len(a), ap, rp,
                   scan_min_range, scan_max_range, thr, C.byref(out))
    return out


def pixel_accurate_score(level: PortPre, sensor_pose, angles, ranges, *, geom: PortMap,
                         score_range_min=0.01, score_range_max=20.0, scan_min_range=0.02,
                         scan_max_range=30.0):
    a, ap = _d(angles)
    r, rp = _d(ranges)
    return lib().orc_pixel_accurate_score(level._dense.ctypes.data_as(c_dp), C.byref(geom.geo),
                                          score_range_min, score_range_max, _arr3(sensor_pose), len(a),
                                          ap, rp, scan_min_range, scan_max_range)


# Names shared with refapi so tests are backend-agnostic.
RefMap, RefPre, RefBuilder = PortMap, PortPre, PortBuilder
