"""ORACLE (test infrastructure): ctypes binding of oracle/_ref/libkarto_ref.so -- the UNMODIFIED
reference karto_sdk compiled by oracle/Makefile plus oracle/ref_driver.cpp."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import REF_DIR

_lib = None
_DP = C.POINTER(C.c_double)


def available() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "libkarto_ref.so"))


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(os.path.join(REF_DIR, "libkarto_ref.so"))
        L.kref_init_laser.restype = C.c_int
        L.kref_init_laser.argtypes = [C.c_double] * 6
        L.kref_laser_set_range_threshold.argtypes = [C.c_double]
        L.kref_mapper_create.restype = C.c_void_p
        L.kref_mapper_destroy.argtypes = [C.c_void_p]
        L.kref_mapper_set.restype = C.c_int
        L.kref_mapper_set.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
        L.kref_matcher_create.restype = C.c_void_p
        L.kref_matcher_create.argtypes = [C.c_void_p] + [C.c_double] * 4
        L.kref_matcher_destroy.argtypes = [C.c_void_p]
        L.kref_scan_create.restype = C.c_void_p
        L.kref_scan_create.argtypes = [_DP, C.c_int, _DP, C.c_int]
        L.kref_scan_destroy.argtypes = [C.c_void_p]
        L.kref_scan_set_pose.argtypes = [C.c_void_p, _DP]
        L.kref_scan_points.restype = C.c_int
        L.kref_scan_points.argtypes = [C.c_void_p, _DP, C.c_int]
        L.kref_scan_sensor_pose.argtypes = [C.c_void_p, _DP]
        L.kref_match.restype = C.c_double
        L.kref_match.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, _DP, _DP]
        L.kref_raster.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.c_int]
        L.kref_correlate.restype = C.c_double
        L.kref_correlate.argtypes = [C.c_void_p, C.c_void_p, _DP, _DP, _DP, C.c_double, C.c_double, C.c_int, C.c_int, _DP, _DP]
        L.kref_grid_info.argtypes = [C.c_void_p, C.POINTER(C.c_int), _DP, _DP]
        L.kref_grid_copy.argtypes = [C.c_void_p, C.POINTER(C.c_uint8)]
        L.kref_kernel_copy.argtypes = [C.c_void_p, C.POINTER(C.c_uint8)]
        L.kref_offsets.restype = C.c_int
        L.kref_offsets.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double, C.POINTER(C.c_int32), C.c_int]
        L.kref_find_valid_points.restype = C.c_int
        L.kref_find_valid_points.argtypes = [C.c_void_p, C.c_void_p, _DP, _DP, C.c_int]
        L.kref_sweep.restype = C.c_double
        L.kref_sweep.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int),
                                 C.c_int, C.c_int, C.c_int, _DP, _DP, _DP]
        L.kref_link_info.argtypes = [_DP] * 5
        L.kref_matrix3_inverse.argtypes = [_DP, _DP]
        L.kref_occupancy_create.restype = C.c_void_p
        L.kref_occupancy_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_double, C.c_int, C.c_double]
        L.kref_occupancy_destroy.argtypes = [C.c_void_p]
        L.kref_occupancy_info.argtypes = [C.c_void_p, C.POINTER(C.c_int), _DP]
        L.kref_occupancy_copy.argtypes = [C.c_void_p, C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(_DP)


def init_laser(min_angle, max_angle, ang_res, min_range, max_range, range_threshold) -> int:
    return lib().kref_init_laser(min_angle, max_angle, ang_res, min_range, max_range, range_threshold)


class RefScan:
    def __init__(self, ranges, pose, uid=0):
        r = np.ascontiguousarray(ranges, dtype=np.float64)
        p = np.ascontiguousarray(pose, dtype=np.float64)
        self.n = len(r)
        self.h = lib().kref_scan_create(_dp(r), len(r), _dp(p), uid)

    def points(self):
        out = np.empty((self.n, 2))
        n = lib().kref_scan_points(self.h, _dp(out), self.n)
        return out[:n]

    def sensor_pose(self):
        out = np.empty(3)
        lib().kref_scan_sensor_pose(self.h, _dp(out))
        return out

    def set_pose(self, pose):
        p = np.ascontiguousarray(pose, dtype=np.float64)
        lib().kref_scan_set_pose(self.h, _dp(p))

    def __del__(self):
        if getattr(self, "h", None):
            lib().kref_scan_destroy(self.h)
            self.h = None


def _ptrs(objs):
    arr = (C.c_void_p * max(1, len(objs)))()
    for i, o in enumerate(objs):
        arr[i] = o.h
    return arr


class RefMapper:
    """A karto::Mapper used only as the parameter holder ScanMatcher reads (Mapper.cpp:590-594, 675-682)."""

    def __init__(self, **params):
        self.h = lib().kref_mapper_create()
        for k, v in params.items():
            if lib().kref_mapper_set(self.h, k.encode(), float(v)) != 0:
                raise KeyError(k)


class RefMatcher:
    def __init__(self, mapper: RefMapper, search_size, resolution, smear, range_threshold):
        self.mapper = mapper
        self.h = lib().kref_matcher_create(mapper.h, search_size, resolution, smear, range_threshold)
        if not self.h:
            raise ValueError("ScanMatcher::Create returned NULL")

    def match(self, query, base, do_penalize=True, do_refine=True):
        mean, cov = np.zeros(3), np.zeros(9)
        r = lib().kref_match(self.h, query.h, _ptrs(base), len(base), int(do_penalize), int(do_refine), _dp(mean), _dp(cov))
        return r, mean, cov.reshape(3, 3)

    def raster(self, query, base):
        lib().kref_raster(self.h, query.h, _ptrs(base), len(base))

    def correlate(self, query, center, sp_off, sp_res, ang_off, ang_res, do_penalize, fine, cov=None):
        mean = np.zeros(3)
        cov = np.zeros(9) if cov is None else np.ascontiguousarray(cov, dtype=np.float64).reshape(9).copy()
        c = np.ascontiguousarray(center, dtype=np.float64)
        o = np.ascontiguousarray(sp_off, dtype=np.float64)
        rs = np.ascontiguousarray(sp_res, dtype=np.float64)
        r = lib().kref_correlate(self.h, query.h, _dp(c), _dp(o), _dp(rs), ang_off, ang_res, int(do_penalize), int(fine),
                                 _dp(mean), _dp(cov))
        return r, mean, cov.reshape(3, 3)

    def grid(self):
        info = (C.c_int * 9)()
        off = np.zeros(2)
        scale = C.c_double()
        lib().kref_grid_info(self.h, info, _dp(off), C.byref(scale))
        info = list(info)
        data = np.empty(info[7], dtype=np.uint8)
        lib().kref_grid_copy(self.h, data.ctypes.data_as(C.POINTER(C.c_uint8)))
        return dict(width=info[0], height=info[1], stride=info[2], roi=(info[3], info[4], info[5], info[6]),
                    data_size=info[7], kernel_size=info[8], offset=(off[0], off[1]), scale=scale.value, data=data)

    def kernel(self):
        k = self.grid()["kernel_size"]
        out = np.empty(k * k, dtype=np.uint8)
        lib().kref_kernel_copy(self.h, out.ctypes.data_as(C.POINTER(C.c_uint8)))
        return out.reshape(k, k)

    def offsets(self, query, angle_center, angle_offset, angle_res):
        n_angles = int(np.floor(angle_offset * 2.0 / angle_res + 0.5)) + 1
        out = np.empty((n_angles, query.n), dtype=np.int32)
        na = lib().kref_offsets(self.h, query.h, angle_center, angle_offset, angle_res,
                                out.ctypes.data_as(C.POINTER(C.c_int32)), query.n)
        assert na == n_angles
        return out

    def find_valid_points(self, scan, viewpoint):
        vp = np.ascontiguousarray(viewpoint, dtype=np.float64)
        out = np.empty((scan.n, 2))
        n = lib().kref_find_valid_points(self.h, scan.h, _dp(vp), _dp(out), scan.n)
        return out[:n].copy()


def sweep(matchers, query, scans, chain_start, do_penalize=False, do_refine=False):
    """Reference MatchScan for every candidate chain on len(matchers) host threads. Returns
    (seconds, resp, mean, cov)."""
    cs = np.ascontiguousarray(chain_start, dtype=np.int32)
    nch = len(cs) - 1
    resp, mean, cov = np.zeros(nch), np.zeros((nch, 3)), np.zeros((nch, 9))
    t = lib().kref_sweep(_ptrs(matchers), len(matchers), query.h, _ptrs(scans), cs.ctypes.data_as(C.POINTER(C.c_int)),
                         nch, int(do_penalize), int(do_refine), _dp(resp), _dp(mean), _dp(cov))
    return t, resp, mean, cov.reshape(nch, 3, 3)


def link_info(p1, p2, cov):
    p1 = np.ascontiguousarray(p1, dtype=np.float64)
    p2 = np.ascontiguousarray(p2, dtype=np.float64)
    c = np.ascontiguousarray(cov, dtype=np.float64).reshape(9)
    d, co = np.zeros(3), np.zeros(9)
    lib().kref_link_info(_dp(p1), _dp(p2), _dp(c), _dp(d), _dp(co))
    return d, co.reshape(3, 3)


def matrix3_inverse(m):
    m = np.ascontiguousarray(m, dtype=np.float64).reshape(9)
    out = np.zeros(9)
    lib().kref_matrix3_inverse(_dp(m), _dp(out))
    return out.reshape(3, 3)


def occupancy(scans, resolution, min_pass_through=-1, occupancy_threshold=-1.0, counters=True):
    """karto::OccupancyGrid::CreateFromScans (Karto.h:5946-5961) on RefScan objects.  Negative parameters =
    the reference's static entry point with its defaults (2, 0.1).  Returns None for an empty scan list, else
    dict(width, height, stride, offset, cells[h, stride] u8, pass / hits [h, stride] u32, seconds)."""
    import time
    t = time.perf_counter()
    h = lib().kref_occupancy_create(_ptrs(scans), len(scans), float(resolution), int(min_pass_through), float(occupancy_threshold))
    dt = time.perf_counter() - t
    if not h:
        return None
    info = (C.c_int * 3)()
    off = np.zeros(2)
    lib().kref_occupancy_info(h, info, _dp(off))
    w, hh, st = info[0], info[1], info[2]
    cells = np.zeros((hh, st), dtype=np.uint8)
    ps = np.zeros((hh, st), dtype=np.uint32) if counters else None
    ht = np.zeros((hh, st), dtype=np.uint32) if counters else None
    lib().kref_occupancy_copy(h, cells.ctypes.data_as(C.POINTER(C.c_uint8)),
                              ps.ctypes.data_as(C.POINTER(C.c_uint32)) if counters else None,
                              ht.ctypes.data_as(C.POINTER(C.c_uint32)) if counters else None)
    lib().kref_occupancy_destroy(h)
    return dict(width=w, height=hh, stride=st, offset=off, cells=cells, passes=ps, hits=ht, seconds=dt)
