"""ctypes binding to oracle/_ref/liblgs_ref.so -- TEST INFRASTRUCTURE ONLY.

liblgs_ref.so is the UNMODIFIED reference hot path (compiled by oracle/Makefile from
/root/reference) behind the flat C wrapper oracle/ref_capi.cpp.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product package never does.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "liblgs_ref.so")

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)

# CostGreedyEndpoint ctor arguments, positionally as slam_launcher.cpp:60-72 passes them
# with launcher_settings_default.json:22-31 (MapResolution .05 -> HitAndMissedDist .075).
DEFAULT_COST = (0.01, 20.0, 0.075, 0.1, 1.0, 0.05, 1.0)


class MatchResult(C.Structure):
    _fields_ = [("found", C.c_int), ("ix", C.c_int), ("iy", C.c_int), ("it", C.c_int),
                ("winX", C.c_int), ("winY", C.c_int), ("winT", C.c_int), ("pad", C.c_int),
                ("stepX", C.c_double), ("stepY", C.c_double), ("stepT", C.c_double),
                ("score", C.c_double), ("sensorPose", C.c_double * 3),
                ("bestSensorPose", C.c_double * 3), ("estPose", C.c_double * 3),
                ("normalizedCost", C.c_double), ("cov", C.c_double * 9)]


def available() -> bool:
    return os.path.exists(LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(f"{LIB_PATH} missing: run `make -C oracle ref` where "
                               "/root/reference exists")
        L = C.CDLL(LIB_PATH)
        vp = C.c_void_p
        L.ref_bresenham.restype = C.c_int
        L.ref_bresenham.argtypes = [C.c_int] * 4 + [c_ip, C.c_int]
        L.ref_bayes_update.restype = C.c_double
        L.ref_bayes_update.argtypes = [C.c_double, C.c_double]
        L.ref_sliding_window_max.argtypes = [c_dp, C.c_int, C.c_int, c_dp]
        for f in (L.ref_compound, L.ref_move_backward, L.ref_inverse_compound):
            f.argtypes = [c_dp, c_dp, c_dp]
        L.ref_map_from_dense.restype = vp
        L.ref_map_from_dense.argtypes = [C.c_double, C.c_int, C.c_int, C.c_int, C.c_double,
                                         C.c_double, c_dp]
        L.ref_map_geometry.argtypes = [vp, c_ip, c_ip, c_dp, c_dp, c_dp]
        L.ref_map_dense.argtypes = [vp, c_dp]
        L.ref_map_destroy.argtypes = [vp]
        L.ref_precompute.restype = vp
        L.ref_precompute.argtypes = [vp, C.c_int]
        L.ref_precompute_pyramid.restype = C.c_int
        L.ref_precompute_pyramid.argtypes = [vp, C.c_int, C.POINTER(vp)]
        L.ref_pre_dense.argtypes = [vp, c_dp]
        L.ref_pre_destroy.argtypes = [vp]
        L.ref_builder_create.restype = vp
        L.ref_builder_create.argtypes = [C.c_double, C.c_int, C.c_int] + [C.c_double] * 5
        scan_args = [c_dp, c_dp, C.c_int, c_dp, c_dp, C.c_double, C.c_double]
        L.ref_builder_append_scan.restype = C.c_int
        L.ref_builder_append_scan.argtypes = [vp] + scan_args
        L.ref_builder_append_node_only.argtypes = [vp] + scan_args
        L.ref_builder_update_grid_map.restype = C.c_int
        L.ref_builder_update_grid_map.argtypes = [vp]
        L.ref_builder_update_latest_map.argtypes = [vp]
        L.ref_builder_construct_map.restype = vp
        L.ref_builder_construct_map.argtypes = [vp, C.c_int, C.c_int]
        L.ref_builder_set_node_pose.argtypes = [vp, C.c_int, c_dp]
        L.ref_builder_after_loop_closure.argtypes = [vp]
        L.ref_builder_num_local_maps.restype = C.c_int
        L.ref_builder_num_local_maps.argtypes = [vp]
        L.ref_builder_local_map.restype = vp
        L.ref_builder_local_map.argtypes = [vp, C.c_int]
        L.ref_builder_local_map_nodes.argtypes = [vp, C.c_int, c_ip, c_ip]
        L.ref_builder_latest_map.restype = vp
        L.ref_builder_latest_map.argtypes = [vp]
        L.ref_builder_destroy.argtypes = [vp]
        L.ref_rtcsm_match.restype = C.c_int
        L.ref_rtcsm_match.argtypes = [vp, vp, C.c_int] + [C.c_double] * 4 + [c_dp] + \
            scan_args + [C.c_double, C.POINTER(MatchResult)]
        L.ref_rtcsm_score_table.restype = C.c_int
        L.ref_rtcsm_score_table.argtypes = [vp, vp, C.c_int, C.c_int, C.c_double, c_dp,
                                            C.c_int, c_dp, c_dp, C.c_double, C.c_int,
                                            C.c_int, C.c_int, C.c_int, C.c_int, c_dp, c_ip,
                                            c_ip]
        L.ref_bb_match.restype = C.c_int
        L.ref_bb_match.argtypes = [vp, C.POINTER(vp), C.c_int] + [C.c_double] * 6 + [c_dp] + \
            scan_args + [C.c_double, C.POINTER(MatchResult)]
        L.ref_pixel_accurate_score.restype = C.c_double
        L.ref_pixel_accurate_score.argtypes = [vp, C.c_double, C.c_double, c_dp, C.c_int,
                                               c_dp, c_dp, C.c_double, C.c_double]
        L.ref_host_tail.argtypes = [vp, c_dp, c_dp, c_dp, C.c_int, c_dp, c_dp, C.c_double,
                                    C.c_double, c_dp, c_dp, c_dp]
        L.ref_cost_greedy_endpoint.restype = C.c_double
        L.ref_cost_greedy_endpoint.argtypes = [vp, c_dp, c_dp, C.c_int, c_dp, c_dp, C.c_double, C.c_double]
        _lib = L
    return _lib


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(c_dp)


def _arr3(p):
    return (C.c_double * 3)(*[float(v) for v in p])


class RefMap:
    """Owns a reference GridMap<BinaryBayesGridCell<double>>."""

    def __init__(self, handle):
        if not handle:
            raise RuntimeError("reference map construction failed")
        self.h = C.c_void_p(handle)

    @classmethod
    def from_dense(cls, dense, min_x, min_y, res=0.05, patch=64):
        dense = np.ascontiguousarray(dense, dtype=np.float64)
        ny, nx = dense.shape
        return cls(lib().ref_map_from_dense(res, patch, nx, ny, min_x, min_y,
                                            dense.ctypes.data_as(c_dp)))

    def geometry(self):
        nx, ny = C.c_int(), C.c_int()
        mx, my, res = C.c_double(), C.c_double(), C.c_double()
        lib().ref_map_geometry(self.h, C.byref(nx), C.byref(ny), C.byref(mx), C.byref(my),
                               C.byref(res))
        return nx.value, ny.value, mx.value, my.value, res.value

    def dense(self):
        nx, ny, *_ = self.geometry()
        out = np.empty((ny, nx), dtype=np.float64)
        lib().ref_map_dense(self.h, out.ctypes.data_as(c_dp))
        return out

    def precompute(self, win):
        return RefPre(lib().ref_precompute(self.h, int(win)), self)

    def pyramid(self, height_max):
        hs = (C.c_void_p * (height_max + 1))()
        lib().ref_precompute_pyramid(self.h, int(height_max), hs)
        return [RefPre(hs[i], self) for i in range(height_max + 1)]

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.ref_map_destroy(self.h)
            self.h = None


class RefPre:
    """Owns a reference GridMap<ConstGridCell<double>> (a win-max map)."""

    def __init__(self, handle, parent: RefMap):
        self.h = C.c_void_p(handle)
        self.shape = parent.geometry()[1::-1]   # (ny, nx)

    def dense(self):
        out = np.empty(self.shape, dtype=np.float64)
        lib().ref_pre_dense(self.h, out.ctypes.data_as(c_dp))
        return out

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.ref_pre_destroy(self.h)
            self.h = None


class RefBuilder:
    """Owns a reference GridMapBuilder + PoseGraph."""

    def __init__(self, res=0.05, patch=64, n_latest=10, travel_thr=20.0, rmin=0.01,
                 rmax=20.0, p_hit=0.6, p_miss=0.45, scan_min_range=0.02,
                 scan_max_range=30.0, rel_pose=(0.0, 0.0, 0.0)):
        self.h = C.c_void_p(lib().ref_builder_create(res, patch, n_latest, travel_thr, rmin,
                                                     rmax, p_hit, p_miss))
        self.scan_min_range, self.scan_max_range = scan_min_range, scan_max_range
        self.rel = _arr3(rel_pose)

    def _scan(self, pose, angles, ranges):
        a, ap = _d(angles)
        r, rp = _d(ranges)
        return (_arr3(pose), self.rel, len(a), ap, rp, self.scan_min_range,
                self.scan_max_range), (a, r)

    def append_scan(self, pose, angles, ranges) -> bool:
        args, keep = self._scan(pose, angles, ranges)
        return bool(lib().ref_builder_append_scan(self.h, *args))

    def append_node_only(self, pose, angles, ranges):
        args, keep = self._scan(pose, angles, ranges)
        lib().ref_builder_append_node_only(self.h, *args)

    def update_grid_map(self) -> bool:
        return bool(lib().ref_builder_update_grid_map(self.h))

    def update_latest_map(self):
        lib().ref_builder_update_latest_map(self.h)

    def construct_map(self, lo, hi) -> RefMap:
        return RefMap(lib().ref_builder_construct_map(self.h, lo, hi))

    def set_node_pose(self, idx, pose):
        lib().ref_builder_set_node_pose(self.h, idx, _arr3(pose))

    def after_loop_closure(self):
        lib().ref_builder_after_loop_closure(self.h)

    def num_local_maps(self):
        return lib().ref_builder_num_local_maps(self.h)

    def local_map(self, i) -> RefMap:
        return RefMap(lib().ref_builder_local_map(self.h, i))

    def local_map_nodes(self, i):
        lo, hi = C.c_int(), C.c_int()
        lib().ref_builder_local_map_nodes(self.h, i, C.byref(lo), C.byref(hi))
        return lo.value, hi.value

    def latest_map(self) -> RefMap:
        return RefMap(lib().ref_builder_latest_map(self.h))

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.ref_builder_destroy(self.h)
            self.h = None


def bresenham(x0, y0, x1, y1):
    cap = abs(x1 - x0) + abs(y1 - y0) + 2
    buf = np.empty((cap, 2), dtype=np.int32)
    n = lib().ref_bresenham(x0, y0, x1, y1, buf.ctypes.data_as(c_ip), cap)
    return buf[:n].copy()


def bayes_update(v, p):
    return lib().ref_bayes_update(float(v), float(p))


def sliding_window_max(a, w):
    a, ap = _d(a)
    out = np.empty_like(a)
    lib().ref_sliding_window_max(ap, len(a), int(w), out.ctypes.data_as(c_dp))
    return out


def compound(a, b):
    o = (C.c_double * 3)()
    lib().ref_compound(_arr3(a), _arr3(b), o)
    return tuple(o)


def move_backward(a, b):
    o = (C.c_double * 3)()
    lib().ref_move_backward(_arr3(a), _arr3(b), o)
    return tuple(o)


def rtcsm_match(m: RefMap, angles, ranges, init_pose, *, low_res=5, range_x=1.0, range_y=1.0,
                range_theta=1.0471975512, scan_range_max=20.0, thr=None, pre: RefPre | None = None,
                rel=(0.0, 0.0, 0.0), scan_min_range=0.02, scan_max_range=30.0,
                cost=DEFAULT_COST) -> MatchResult:
    """ScanMatcherRealTimeCorrelative::OptimizePose (5-argument overload).

    thr=None passes std::numeric_limits<double>::min() like the 1-argument overload does
    (scan_matcher_real_time_correlative.cpp:45-47)."""
    a, ap = _d(angles)
    r, rp = _d(ranges)
    out = MatchResult()
    costv = (C.c_double * 7)(*cost)
    if thr is None:
        thr = float(np.finfo(np.float64).tiny)
    lib().ref_rtcsm_match(m.h, pre.h if pre is not None else None, low_res, range_x, range_y,
                          range_theta, scan_range_max, costv, _arr3(init_pose), _arr3(rel), len(a),
                          ap, rp, scan_min_range, scan_max_range, thr, C.byref(out))
    return out


def rtcsm_score_table(m: RefMap, pre: RefPre, use_coarse, low_res, scan_range_max,
                      sensor_pose, angles, ranges, step_t, win_t, x_lo, nxw, y_lo, nyw,
                      want_table=True):
    a, ap = _d(angles)
    r, rp = _d(ranges)
    nt = 2 * win_t + 1
    table = np.empty((nt, nyw, nxw), dtype=np.float64) if want_table else None
    idx = np.full((nt, len(a), 2), -(2 ** 31), dtype=np.int32)
    cnt = np.zeros(nt, dtype=np.int32)
    lib().ref_rtcsm_score_table(m.h, pre.h, int(use_coarse), low_res, scan_range_max,
                                _arr3(sensor_pose), len(a), ap, rp, step_t, win_t, x_lo, nxw,
                                y_lo, nyw,
                                table.ctypes.data_as(c_dp) if want_table else None,
                                idx.ctypes.data_as(c_ip), cnt.ctypes.data_as(c_ip))
    return table, idx, cnt


def bb_match(m: RefMap, angles, ranges, init_pose, *, height_max=6, range_x=2.0, range_y=2.0,
             range_theta=1.0, scan_range_max=20.0, score_range_min=0.01, score_range_max=20.0,
             thr=0.6, pyramid=None, rel=(0.0, 0.0, 0.0), scan_min_range=0.02,
             scan_max_range=30.0, cost=DEFAULT_COST) -> MatchResult:
    """ScanMatcherBranchBound::OptimizePose (5-argument overload)."""
    a, ap = _d(angles)
    r, rp = _d(ranges)
    out = MatchResult()
    costv = (C.c_double * 7)(*cost)
    pyr = None
    if pyramid is not None:
        pyr = (C.c_void_p * len(pyramid))(*[p.h for p in pyramid])
    lib().ref_bb_match(m.h, pyr, height_max, range_x, range_y, range_theta, scan_range_max,
                       score_range_min, score_range_max, costv, _arr3(init_pose), _arr3(rel),
                       len(a), ap, rp, scan_min_range, scan_max_range, thr, C.byref(out))
    return out


def gs_match(m: RefMap, angles, ranges, init_pose, *, range_x=2.0, range_y=2.0, range_theta=0.5,
             step_x=0.05, step_y=0.05, step_theta=0.005, score_range_min=0.01, score_range_max=20.0,
             thr=0.5, rel=(0.0, 0.0, 0.0), scan_min_range=0.02, scan_max_range=30.0,
             cost=DEFAULT_COST) -> MatchResult:
    """ScanMatcherGridSearch::OptimizePose (4-argument overload); ix / iy / it = winning loop counters,
    winX / winY / winT = loop lengths."""
    a, ap = _d(angles)
    r, rp = _d(ranges)
    out = MatchResult()
    L = lib()
    L.ref_gs_match.argtypes = [C.c_void_p] + [C.c_double] * 8 + [c_dp, c_dp, c_dp, C.c_int, c_dp, c_dp,
                                                                  C.c_double, C.c_double, C.c_double,
                                                                  C.POINTER(MatchResult)]
    L.ref_gs_match(m.h, range_x, range_y, range_theta, step_x, step_y, step_theta, score_range_min,
                   score_range_max, (C.c_double * 7)(*cost), _arr3(init_pose), _arr3(rel), len(a), ap, rp,
                   scan_min_range, scan_max_range, thr, C.byref(out))
    return out


def pixel_accurate_score(level: RefPre, sensor_pose, angles, ranges, *, score_range_min=0.01,
                         score_range_max=20.0, scan_min_range=0.02, scan_max_range=30.0):
    a, ap = _d(angles)
    r, rp = _d(ranges)
    return lib().ref_pixel_accurate_score(level.h, score_range_min, score_range_max,
                                          _arr3(sensor_pose), len(a), ap, rp, scan_min_range,
                                          scan_max_range)


def host_tail(m: RefMap, best_sensor_pose, angles, ranges, *, rel=(0.0, 0.0, 0.0),
              scan_min_range=0.02, scan_max_range=30.0, cost=DEFAULT_COST):
    a, ap = _d(angles)
    r, rp = _d(ranges)
    nc = C.c_double()
    est = (C.c_double * 3)()
    cov = (C.c_double * 9)()
    lib().ref_host_tail(m.h, (C.c_double * 7)(*cost), _arr3(best_sensor_pose), _arr3(rel),
                        len(a), ap, rp, scan_min_range, scan_max_range, C.byref(nc), est, cov)
    return nc.value, tuple(est), np.array(cov).reshape(3, 3)


def cost_greedy_endpoint(m: RefMap, sensor_pose, angles, ranges, *, scan_min_range=0.02,
                         scan_max_range=30.0, cost=DEFAULT_COST) -> float:
    """CostGreedyEndpoint::Cost at one sensor pose."""
    a, ap = _d(angles)
    r, rp = _d(ranges)
    return lib().ref_cost_greedy_endpoint(m.h, (C.c_double * 7)(*cost), _arr3(sensor_pose), len(a), ap, rp,
                                          scan_min_range, scan_max_range)


# ---- integration helpers ---------------------------------------------------------------------
def _bind_integration():
    L = lib()
    if getattr(L, "_integ_bound", False):
        return L
    vp = C.c_void_p
    L.ref_hit_points.restype = C.c_int
    L.ref_hit_points.argtypes = [c_dp, c_dp, C.c_int, c_dp, c_dp] + [C.c_double] * 4 + [c_dp] * 3
    L.ref_map_integrate_hits.restype = C.c_int
    L.ref_map_integrate_hits.argtypes = [vp, c_dp, C.c_int, c_dp, C.c_double, C.c_double]
    L.ref_map_resize.argtypes = [vp] + [C.c_double] * 4
    L.ref_map_expand.argtypes = [vp] + [C.c_double] * 5
    L.ref_map_reset.argtypes = [vp]
    L.ref_map_create_empty.restype = vp
    L.ref_map_create_empty.argtypes = [C.c_double, C.c_int, C.c_double, C.c_double]
    L._integ_bound = True
    return L


def hit_points(robot_pose, angles, ranges, *, rel=(0.0, 0.0, 0.0), scan_min_range=0.02,
               scan_max_range=30.0, usable_min=0.01, usable_max=20.0):
    """-> (sensor_pose[3], hit_xy[k][2], bbox[4]) via ComputeBoundingBoxAndScanPoints."""
    L = _bind_integration()
    a, ap = _d(angles)
    r, rp = _d(ranges)
    sp = (C.c_double * 3)()
    bbox = (C.c_double * 4)()
    out = np.empty((len(a), 2), dtype=np.float64)
    k = L.ref_hit_points(_arr3(robot_pose), _arr3(rel), len(a), ap, rp, scan_min_range,
                         scan_max_range, usable_min, usable_max, sp, out.ctypes.data_as(c_dp), bbox)
    return np.array(sp), out[:k].copy(), np.array(bbox)


def empty_map(center_x, center_y, res=0.05, patch=64) -> RefMap:
    return RefMap(_bind_integration().ref_map_create_empty(res, patch, center_x, center_y))


def map_integrate_hits(m: RefMap, sensor_xy, hit_xy, p_hit=0.6, p_miss=0.45) -> int:
    L = _bind_integration()
    s, sp = _d(np.asarray(sensor_xy, dtype=np.float64)[:2])
    h, hp = _d(hit_xy)
    n = L.ref_map_integrate_hits(m.h, sp, len(h), hp, p_hit, p_miss)
    if n < 0:
        raise RuntimeError("a touched cell lies outside the oracle map")
    return n


def map_resize(m: RefMap, min_x, min_y, max_x, max_y):
    _bind_integration().ref_map_resize(m.h, min_x, min_y, max_x, max_y)


def map_expand(m: RefMap, min_x, min_y, max_x, max_y, step=5.0):
    _bind_integration().ref_map_expand(m.h, min_x, min_y, max_x, max_y, step)


def map_reset(m: RefMap):
    _bind_integration().ref_map_reset(m.h)
