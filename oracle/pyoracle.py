"""ctypes binding of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU arms
(cpu_baseline / --impl reference).  The product package `cover_b200` never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

SHAPE_OUTLINE, SHAPE_AREA = 1, 2
OP_IS_FREE, OP_ACCUMULATE, OP_MAX = 0, 1, 2
POSE_CS, POSE_LATTICE, POSE_NONE = 0, 1, 2
OK, LOGIC_ERROR, ROW_COMPLEXITY, COORD_RANGE = 0, 1, 2, 3


class _Map(C.Structure):
    _fields_ = [("data", C.c_void_p), ("size_x", C.c_int32), ("size_y", C.c_int32),
                ("resolution", C.c_double), ("origin_x", C.c_double), ("origin_y", C.c_double)]


class _Lattice(C.Structure):
    _fields_ = [("x0", C.c_double), ("dx", C.c_double), ("y0", C.c_double), ("dy", C.c_double),
                ("cs", C.c_void_p), ("nx", C.c_int32), ("ny", C.c_int32), ("ntheta", C.c_int32)]


def build(force: bool = False) -> str:
    """Compile liboracle.so (and oracle/_ref when the reference tree is present)."""
    src_newer = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB_PATH)
        for f in ("oracle_capi.cpp", "cover_oracle.hpp"))
    if force or src_newer:
        subprocess.run(["make", "-C", _HERE, "liboracle.so"], check=True, capture_output=True)
    return _LIB_PATH


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


@dataclass
class Costmap:
    """uint8 row-major costmap [size_y, size_x] + geometry (stand-in for costmap_2d::Costmap2D)."""
    data: np.ndarray
    resolution: float
    origin_x: float = 0.0
    origin_y: float = 0.0

    def __post_init__(self):
        self.data = np.ascontiguousarray(self.data, dtype=np.uint8)
        assert self.data.ndim == 2

    @property
    def size_x(self):
        return self.data.shape[1]

    @property
    def size_y(self):
        return self.data.shape[0]

    def c(self):
        return _Map(self.data.ctypes.data, self.size_x, self.size_y, self.resolution, self.origin_x, self.origin_y)


@dataclass
class Lattice:
    """Pose lattice: index p = (it*ny + iy)*nx + ix, t = (x0 + ix*dx, y0 + iy*dy), heading table cs[it]."""
    x0: float
    dx: float
    nx: int
    y0: float
    dy: float
    ny: int
    cs: np.ndarray  # [ntheta, 2] = (cos, sin) evaluated on the host

    def __post_init__(self):
        self.cs = np.ascontiguousarray(self.cs, dtype=np.float64).reshape(-1, 2)

    @property
    def count(self):
        return self.nx * self.ny * self.cs.shape[0]

    def c(self):
        return _Lattice(self.x0, self.dx, self.y0, self.dy, self.cs.ctypes.data, self.nx, self.ny, self.cs.shape[0])

    def poses_cs(self, first=0, count=None):
        """Materialise poses [count, 4] = (c, s, tx, ty) with the same arithmetic as the lattice rule."""
        count = self.count - first if count is None else count
        p = np.arange(first, first + count, dtype=np.int64)
        ix, iy, it = p % self.nx, (p // self.nx) % self.ny, p // (self.nx * self.ny)
        out = np.empty((count, 4))
        out[:, 0:2] = self.cs[it]
        out[:, 2] = self.x0 + ix.astype(np.float64) * self.dx
        out[:, 3] = self.y0 + iy.astype(np.float64) * self.dy
        return out


@dataclass
class Result:
    is_free: np.ndarray
    cost: np.ndarray
    status: np.ndarray
    cells_read: int


def polygon_xy(poly) -> np.ndarray:
    """Accepts [2, V] (Eigen-like rows x/y) and returns the column-major flat x0,y0,x1,y1,... buffer."""
    a = np.asarray(poly, dtype=np.float64)
    assert a.ndim == 2 and a.shape[0] == 2
    return np.ascontiguousarray(a.T).reshape(-1)


def hardware_threads() -> int:
    return int(default()._lib.oracle_hardware_threads())


def _outputs(n, want_cost=True):
    return (np.zeros(n, dtype=np.uint8), np.zeros(n, dtype=np.float64) if want_cost else None, np.zeros(n, dtype=np.uint8))


def _pose_args(poses):
    if poses is None:
        return POSE_NONE, None, None, 1, None
    if isinstance(poses, Lattice):
        lat = poses.c()
        return POSE_LATTICE, None, C.byref(lat), poses.count, lat
    arr = np.ascontiguousarray(poses, dtype=np.float64).reshape(-1, 4)
    return POSE_CS, arr, None, len(arr), arr


def pack_polygons(polys):
    """list of [2, V_k] -> (flat column-major xy, starts[int32, K+1])."""
    starts = np.zeros(len(polys) + 1, dtype=np.int32)
    flat = []
    for k, p in enumerate(polys):
        xy = polygon_xy(p)
        flat.append(xy)
        starts[k + 1] = starts[k] + xy.size // 2
    xy = np.concatenate(flat) if flat else np.zeros(0)
    return np.ascontiguousarray(xy, dtype=np.float64), starts




class Backend:
    """One loaded checker library: the restatement (liboracle.so) or the reference itself (oracle/_ref)."""

    def __init__(self, path, name):
        self.path, self.name = path, name
        self._lib = C.CDLL(path)
        self._lib.oracle_ray_cells.restype = C.c_int64
        self._lib.oracle_outline_cells.restype = C.c_int64
        self._lib.oracle_area_cells.restype = C.c_int64

    def to_discrete(self, res, poly) -> np.ndarray:
        xy = polygon_xy(poly)
        n = xy.size // 2
        out = np.empty((n, 2), dtype=np.int32)
        rc = self._lib.oracle_to_discrete(C.c_double(res), _p(xy), C.c_int64(n), _p(out))
        if rc != 0:
            raise ValueError("The resolution must be positive")
        return out



    def ray_cells(self, b, e) -> np.ndarray:
        cap = max(abs(e[0] - b[0]), abs(e[1] - b[1])) + 1
        out = np.empty((cap, 2), dtype=np.int32)
        n = self._lib.oracle_ray_cells(int(b[0]), int(b[1]), int(e[0]), int(e[1]), _p(out), C.c_int64(cap))
        return out[:n]



    def outline_cells(self, sparse) -> np.ndarray:
        s = np.ascontiguousarray(sparse, dtype=np.int32).reshape(-1, 2)
        n = self._lib.oracle_outline_cells(_p(s), C.c_int64(len(s)), None, C.c_int64(0))
        out = np.empty((max(n, 1), 2), dtype=np.int32)
        self._lib.oracle_outline_cells(_p(s), C.c_int64(len(s)), _p(out), C.c_int64(n))
        return out[:n]



    def area_cells(self, dense_outlines, return_max_ranges=False):
        """dense_outlines: list of [n_i, 2] int arrays (each a dense closed outline). Raises RuntimeError for
        the reference's std::logic_error."""
        outs = [np.ascontiguousarray(o, dtype=np.int32).reshape(-1, 2) for o in dense_outlines]
        sizes = np.array([len(o) for o in outs], dtype=np.int64)
        cat = np.concatenate(outs) if outs else np.empty((0, 2), dtype=np.int32)
        cat = np.ascontiguousarray(cat)
        mr = C.c_int32(0)
        n = self._lib.oracle_area_cells(_p(cat), _p(sizes), C.c_int32(len(outs)), None, C.c_int64(0), C.byref(mr))
        if n < 0:
            raise RuntimeError("Even number of line-pairs expected")
        out = np.empty((max(n, 1), 2), dtype=np.int32)
        self._lib.oracle_area_cells(_p(cat), _p(sizes), C.c_int32(len(outs)), _p(out), C.c_int64(n), C.byref(mr))
        return (out[:n], mr.value) if return_max_ranges else out[:n]



    def polygon_batch(self, cmap: Costmap, poly, poses, shape=SHAPE_AREA, op=OP_IS_FREE, first=0, count=None, threads=1, offset=None) -> Result:
        """`offset` (2 ints): the offset overloads is_free / accumulate_cost(gen, offset, map), impl/costmap.hpp:29-35,74-81."""
        xy = polygon_xy(poly)
        fmt, arr, lat, total, _keep = _pose_args(poses)
        n = total - first if count is None else count
        free, cost, status = _outputs(n)
        read = C.c_int64(0)
        m = cmap.c()
        if offset is not None:
            off = np.ascontiguousarray(offset, dtype=np.int32).reshape(2)
            rc = self._lib.oracle_polygon_batch_offset(C.byref(m), _p(xy), C.c_int32(xy.size // 2), C.c_int32(shape), C.c_int32(op), C.c_int32(fmt), _p(arr),
                                                       lat, C.c_int64(first), C.c_int64(n), _p(off), _p(free), _p(cost), _p(status), C.byref(read),
                                                       C.c_int32(threads))
            if rc != 0:
                raise ValueError("oracle_polygon_batch_offset failed (resolution must be positive)")
            return Result(free, cost, status, read.value)
        rc = self._lib.oracle_polygon_batch(C.byref(m), _p(xy), C.c_int32(xy.size // 2), C.c_int32(shape), C.c_int32(op),
                                        C.c_int32(fmt), _p(arr), lat, C.c_int64(first), C.c_int64(n), _p(free), _p(cost),
                                        _p(status), C.byref(read), C.c_int32(threads))
        if rc != 0:
            raise ValueError("oracle_polygon_batch failed (resolution must be positive)")
        return Result(free, cost, status, read.value)



    def multi_polygon_batch(self, cmap: Costmap, outlines, poses, op=OP_IS_FREE, first=0, count=None, threads=1) -> Result:
        """area_generator(res, polygon_vec) checks (holes): src/cover/generators.cpp:126-135."""
        xy, starts = pack_polygons(outlines)
        fmt, arr, lat, total, _keep = _pose_args(poses)
        n = total - first if count is None else count
        free, cost, status = _outputs(n)
        read = C.c_int64(0)
        m = cmap.c()
        rc = self._lib.oracle_multi_polygon_batch(C.byref(m), _p(xy), _p(starts), C.c_int32(len(outlines)), C.c_int32(op), C.c_int32(fmt), _p(arr),
                                                   lat, C.c_int64(first), C.c_int64(n), _p(free), _p(cost), _p(status), C.byref(read), C.c_int32(threads))
        if rc != 0:
            raise ValueError("oracle_multi_polygon_batch failed")
        return Result(free, cost, status, read.value)

    def rays_batch(self, cmap: Costmap, segments, offset=None, op=OP_IS_FREE, threads=1) -> Result:
        seg = np.ascontiguousarray(segments, dtype=np.int32).reshape(-1, 4)
        off = None if offset is None else np.ascontiguousarray(offset, dtype=np.int32)
        free, cost, status = _outputs(len(seg))
        read = C.c_int64(0)
        m = cmap.c()
        self._lib.oracle_rays_batch(C.byref(m), _p(seg), C.c_int64(len(seg)), _p(off), C.c_int32(op), _p(free), _p(cost),
                                _p(status), C.byref(read), C.c_int32(threads))
        return Result(free, cost, status, read.value)



    def cells_batch(self, cmap: Costmap, cells, offsets, op=OP_IS_FREE, threads=1) -> Result:
        cl = np.ascontiguousarray(cells, dtype=np.int32).reshape(-1, 2)
        off = np.ascontiguousarray(offsets, dtype=np.int32).reshape(-1, 2)
        free, cost, status = _outputs(len(off))
        read = C.c_int64(0)
        m = cmap.c()
        self._lib.oracle_cells_batch(C.byref(m), _p(cl), C.c_int64(len(cl)), _p(off), C.c_int64(len(off)), C.c_int32(op),
                                 _p(free), _p(cost), _p(status), C.byref(read), C.c_int32(threads))
        return Result(free, cost, status, read.value)



    def make_discrete_footprint(self, res, outlines, areas):
        """generator_footprint::init + discrete_footprint::init: returns (outline_cells[No,2], area_cells[Na,2])."""
        oxy, ost = pack_polygons(outlines)
        axy, ast = pack_polygons(areas)
        no, na = C.c_int64(0), C.c_int64(0)
        args = (C.c_double(res), _p(oxy), _p(ost), C.c_int32(len(outlines)), _p(axy), _p(ast), C.c_int32(len(areas)))
        rc = self._lib.oracle_make_discrete_footprint(*args, None, C.byref(no), None, C.byref(na))
        if rc == -2:
            raise RuntimeError("Even number of line-pairs expected")
        if rc != 0:
            raise ValueError("The resolution must be positive")
        oc = np.empty((max(no.value, 1), 2), dtype=np.int32)
        ac = np.empty((max(na.value, 1), 2), dtype=np.int32)
        self._lib.oracle_make_discrete_footprint(*args, _p(oc), C.byref(no), _p(ac), C.byref(na))
        return oc[:no.value], ac[:na.value]



    def discrete_footprint_batch(self, cmap: Costmap, bank, offsets, which=None, threads=1) -> Result:
        """bank: list of (outline_cells[No,2], area_cells[Na,2]); which[i] selects the footprint of pose i."""
        o_starts = np.zeros(len(bank) + 1, dtype=np.int64)
        a_starts = np.zeros(len(bank) + 1, dtype=np.int64)
        for k, (o, a) in enumerate(bank):
            o_starts[k + 1] = o_starts[k] + len(o)
            a_starts[k + 1] = a_starts[k] + len(a)
        oc = np.ascontiguousarray(np.concatenate([np.asarray(o, dtype=np.int32).reshape(-1, 2) for o, _ in bank]))
        ac = np.ascontiguousarray(np.concatenate([np.asarray(a, dtype=np.int32).reshape(-1, 2) for _, a in bank]))
        off = np.ascontiguousarray(offsets, dtype=np.int32).reshape(-1, 2)
        wh = None if which is None else np.ascontiguousarray(which, dtype=np.int32)
        free, _, status = _outputs(len(off), want_cost=False)
        read = C.c_int64(0)
        m = cmap.c()
        self._lib.oracle_discrete_footprint_batch(C.byref(m), _p(oc), _p(o_starts), _p(ac), _p(a_starts), C.c_int32(len(bank)),
                                              _p(off), _p(wh), C.c_int64(len(off)), _p(free), _p(status), C.byref(read),
                                              C.c_int32(threads))
        return Result(free, None, status, read.value)



    def continuous_footprint_batch(self, cmap: Costmap, outlines, areas, poses, first=0, count=None, threads=1) -> Result:
        oxy, ost = pack_polygons(outlines)
        axy, ast = pack_polygons(areas)
        fmt, arr, lat, total, _keep = _pose_args(poses)
        n = total - first if count is None else count
        free, _, status = _outputs(n, want_cost=False)
        read = C.c_int64(0)
        m = cmap.c()
        rc = self._lib.oracle_continuous_footprint_batch(C.byref(m), _p(oxy), _p(ost), C.c_int32(len(outlines)), _p(axy), _p(ast),
                                                     C.c_int32(len(areas)), C.c_int32(fmt), _p(arr), lat, C.c_int64(first),
                                                     C.c_int64(n), _p(free), _p(status), C.byref(read), C.c_int32(threads))
        if rc != 0:
            raise ValueError("oracle_continuous_footprint_batch failed")
        return Result(free, None, status, read.value)



_default = None
_reference = None
REF_LIB_PATH = os.path.join(_HERE, "_ref", "libcover_ref.so")


def default() -> Backend:
    """The CPU restatement (always available: built on demand with g++)."""
    global _default
    if _default is None:
        _default = Backend(build(), "port")
    return _default


def reference():
    """The reference's own sources compiled over shim headers, or None when oracle/_ref was not built
    (it is built in the authoring container where /root/reference exists and travels with gpurun)."""
    global _reference
    if _reference is None and os.path.exists(REF_LIB_PATH):
        _reference = Backend(REF_LIB_PATH, "reference")
    return _reference


def _delegate(name):
    def f(*a, **k):
        return getattr(default(), name)(*a, **k)
    f.__name__ = name
    return f


for _n in ['to_discrete', 'ray_cells', 'outline_cells', 'area_cells', 'polygon_batch', 'rays_batch', 'cells_batch', 'make_discrete_footprint', 'discrete_footprint_batch', 'continuous_footprint_batch', 'multi_polygon_batch']:
    globals()[_n] = _delegate(_n)
