"""cover_b200 — B200-native batched footprint collision checks (many poses x one costmap).

Python host layer over the C ABI in include/cover_b200.h (ctypes; no torch dependency).  It mirrors the
reference's surface for this path — `is_free` / `accumulate_cost` over ray / outline / area generators
(src/cover/costmap.hpp:190-306), cached `discrete_polygon`s (costmap.cpp:59-69) and the three split
footprint classes (src/cover/footprints.hpp:117-252) — with a leading batch dimension.

There is NO CPU fallback: every call goes through libcover_b200.so into CUDA kernels and raises
`CoverError` when no CUDA device (or no built library) is available.

Array conventions: polygons are `[2, V]` float64 (row 0 = x, row 1 = y, like Eigen's cover::polygon);
cell lists `[N, 2]` int32; poses `[N, 4]` float64 = (cos, sin, tx, ty) with the trigonometry evaluated on
the host.  Host inputs are numpy arrays; device inputs/outputs are anything with `.data_ptr()` and
`.is_cuda` (torch tensors) — they are used in place, without copies.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from . import _build

__all__ = [
    "CoverError", "Context", "Costmap", "Lattice", "Result", "DiscretePolygon", "FootprintBank", "SplitFootprint",
    "trajectory_reduce", "get_radius", "split_polygon", "STATUS_OK", "STATUS_LOGIC_ERROR", "STATUS_ROW_COMPLEXITY", "STATUS_COORD_RANGE", "library_path", "build",
]

STATUS_OK, STATUS_LOGIC_ERROR, STATUS_ROW_COMPLEXITY, STATUS_COORD_RANGE = 0, 1, 2, 3
_MEM_HOST, _MEM_DEVICE, _MEM_HOST_ASYNC = 0, 1, 2
_POSE_CS, _POSE_LATTICE, _POSE_NONE = 0, 1, 2
_ERRORS = {-1: "invalid argument", -2: "CUDA error", -3: "out of memory", -4: "unsupported"}


class CoverError(RuntimeError):
    """A negative cover_error from the C ABI.  `code == -1` corresponds to the reference's
    std::invalid_argument (e.g. non-positive resolution, base.hpp:54-55)."""

    def __init__(self, code: int, message: str):
        super().__init__(f"cover_b200: {_ERRORS.get(code, 'error')} ({code}): {message}")
        self.code = code


class _Lattice(C.Structure):
    _fields_ = [("x0", C.c_double), ("dx", C.c_double), ("y0", C.c_double), ("dy", C.c_double), ("cs", C.c_void_p),
                ("nx", C.c_int32), ("ny", C.c_int32), ("ntheta", C.c_int32)]


class _Poses(C.Structure):
    _fields_ = [("format", C.c_int32), ("mem", C.c_int32), ("first", C.c_int64), ("count", C.c_int64), ("data", C.c_void_p),
                ("lattice", _Lattice), ("cell_offset", C.c_int32 * 2)]


class _Results(C.Structure):
    _fields_ = [("mem", C.c_int32), ("is_free", C.c_void_p), ("cost", C.c_void_p), ("status", C.c_void_p),
                ("free_bits", C.c_void_p), ("not_ok", C.c_void_p)]


_lib = None


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA library in-tree (nvcc -gencode arch=compute_100a,code=sm_100a)."""
    return _build.build(force=force, verbose=verbose)


def library_path() -> str:
    return _build.LIB_PATH


def _load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_build.LIB_PATH):
        _build.build()  # raises when nvcc is absent: there is no other implementation to fall back to
    lib = C.CDLL(_build.LIB_PATH)
    lib.cover_last_error.restype = C.c_char_p
    lib.cover_ctx_stream.restype = C.c_void_p
    lib.cover_ctx_launch_count.restype = C.c_int64
    for name in ("cover_ctx_destroy", "cover_map_destroy", "cover_cells_destroy", "cover_footprints_destroy", "cover_split_destroy"):
        getattr(lib, name).restype = None
    if lib.cover_abi_version() != 2:
        raise CoverError(-4, "libcover_b200.so has a different ABI version; rebuild it")
    _lib = lib
    return lib


def _is_device(a) -> bool:
    return hasattr(a, "data_ptr") and bool(getattr(a, "is_cuda", False))


def _ptr(a) -> Optional[int]:
    if a is None:
        return None
    if hasattr(a, "data_ptr"):
        return int(a.data_ptr())
    return int(a.ctypes.data)


def polygon_xy(polygon) -> np.ndarray:
    """[2, V] -> flat column-major x0,y0,x1,y1,... (the memory image of an Eigen 2xN matrix)."""
    a = np.asarray(polygon, dtype=np.float64)
    if a.ndim != 2 or a.shape[0] != 2:
        raise ValueError("polygon must have shape [2, V]")
    return np.ascontiguousarray(a.T).reshape(-1)


@dataclass
class Lattice:
    """(x, y, heading) pose lattice expanded on the device.  Pose index p = (it*ny + iy)*nx + ix;
    translation (x0 + ix*dx, y0 + iy*dy) (multiply, then add); heading (cos, sin) = cs[it]."""
    x0: float
    dx: float
    nx: int
    y0: float
    dy: float
    ny: int
    cs: np.ndarray

    def __post_init__(self):
        self.cs = np.ascontiguousarray(self.cs, dtype=np.float64).reshape(-1, 2)

    @property
    def count(self) -> int:
        return int(self.nx) * int(self.ny) * int(self.cs.shape[0])


@dataclass
class Result:
    """Caller-owned outputs.  Any field may be None.  `free_bits` is the bit-packed verdict array
    (uint32 words, bit i & 31 of word i >> 5); `not_ok` a one-element int64 array counting poses whose
    status is not STATUS_OK."""
    is_free: Optional[object] = None
    cost: Optional[object] = None
    status: Optional[object] = None
    free_bits: Optional[object] = None
    not_ok: Optional[object] = None
    host_async: bool = False  # host arrays (page-locked) filled without waiting: valid after Context.synchronize()

    def unpack_bits(self, count: int) -> np.ndarray:
        """free_bits -> uint8 verdicts [count] (host arrays only)."""
        words = np.asarray(self.free_bits, dtype=np.uint32)
        return np.unpackbits(words.view(np.uint8), bitorder="little")[:count]


class Context:
    """One CUDA device + stream + scratch memory (cover_ctx).  Single-threaded."""

    def __init__(self, device: int = 0):
        self._lib = _load()
        h = C.c_void_p()
        rc = self._lib.cover_ctx_create(C.c_int(device), C.byref(h))
        if rc != 0:
            raise CoverError(rc, "cover_ctx_create failed: no usable CUDA device (cover_b200 has no CPU fallback)")
        self._h = h
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            self._lib.cover_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def stream(self) -> int:
        """The cudaStream_t (as an int) every kernel of this context is launched on."""
        return int(self._lib.cover_ctx_stream(self._h))

    @property
    def launch_count(self) -> int:
        return int(self._lib.cover_ctx_launch_count(self._h))

    def synchronize(self):
        self._check(self._lib.cover_ctx_synchronize(self._h))

    def bind(self):
        """Make this context's device the calling thread's current CUDA device (call once in a thread that drives
        this context; see cover_ctx_bind)."""
        self._check(self._lib.cover_ctx_bind(self._h))

    def probe_read_bandwidth(self, nbytes: int, reps: int = 20) -> float:
        """Best-of-10 streaming read bandwidth in GB/s over a scratch buffer (16 MB: L2 -> SM path; 1 GB: HBM)."""
        gbs = C.c_double(0.0)
        self._check(self._lib.cover_probe_read_bandwidth(self._h, C.c_int64(nbytes), C.c_int32(reps), C.byref(gbs)))
        return float(gbs.value)

    def _check(self, rc: int):
        if rc != 0:
            raise CoverError(rc, self._lib.cover_last_error(self._h).decode())

    # -- argument marshalling ------------------------------------------------------------------
    def _poses(self, poses, first: int, count: Optional[int], offset=None):
        keep = []
        p = _Poses()
        if offset is not None:  # the offset overloads: added to every generated cell (impl/costmap.hpp:29-35,74-81)
            p.cell_offset[0], p.cell_offset[1] = int(offset[0]), int(offset[1])
        if poses is None:
            p.format, p.mem, p.first, p.count, p.data = _POSE_NONE, _MEM_HOST, 0, 1, None
        elif isinstance(poses, Lattice):
            total = poses.count
            n = total - first if count is None else count
            p.format, p.mem, p.first, p.count, p.data = _POSE_LATTICE, _MEM_HOST, first, n, None
            p.lattice = _Lattice(poses.x0, poses.dx, poses.y0, poses.dy, poses.cs.ctypes.data, poses.nx, poses.ny, poses.cs.shape[0])
            keep.append(poses.cs)
        elif _is_device(poses):
            if poses.dim() != 2 or poses.shape[1] != 4 or not poses.is_contiguous() or poses.element_size() != 8:
                raise ValueError("device poses must be a contiguous [N, 4] float64 tensor")
            p.format, p.mem, p.first, p.count, p.data = _POSE_CS, _MEM_DEVICE, 0, int(poses.shape[0]), _ptr(poses)
            keep.append(poses)
        else:
            arr = np.ascontiguousarray(poses, dtype=np.float64).reshape(-1, 4)
            p.format, p.mem, p.first, p.count, p.data = _POSE_CS, _MEM_HOST, 0, len(arr), arr.ctypes.data
            keep.append(arr)
        return p, keep

    def _results(self, n: int, want_free: bool, want_cost: bool, want_status: bool, out: Optional[Result]):
        if out is not None:
            arrays = [a for a in (out.is_free, out.cost, out.status, out.free_bits, out.not_ok) if a is not None]
            dev = any(_is_device(a) for a in arrays)
            if dev and not all(_is_device(a) for a in arrays):
                raise ValueError("result arrays must all be host or all be device memory")
            r = _Results(_MEM_DEVICE if dev else (_MEM_HOST_ASYNC if out.host_async else _MEM_HOST), _ptr(out.is_free), _ptr(out.cost), _ptr(out.status),
                         _ptr(out.free_bits), _ptr(out.not_ok))
            return r, out
        res = Result(np.empty(n, dtype=np.uint8) if want_free else None, np.empty(n, dtype=np.float64) if want_cost else None,
                     np.empty(n, dtype=np.uint8) if want_status else None)
        return _Results(_MEM_HOST, _ptr(res.is_free), _ptr(res.cost), _ptr(res.status), None, None), res


def trajectory_reduce(ctx: "Context", is_free=None, cost=None, steps: int = 1):
    """Scores trajectories of `steps` consecutive poses each (extension; see cover_trajectory_reduce).
    Host (numpy) or device (torch) arrays; returns (first_blocked int32 [T] or None, score float64 [T] or None)."""
    ref = is_free if is_free is not None else cost
    n = int(ref.shape[0])
    if n % steps:
        raise ValueError("the number of poses must be a multiple of steps")
    t = n // steps
    dev = _is_device(ref)
    if dev:
        import torch
        first = torch.empty(t, dtype=torch.int32, device=ref.device) if is_free is not None else None
        score = torch.empty(t, dtype=torch.float64, device=ref.device) if cost is not None else None
    else:
        is_free = None if is_free is None else np.ascontiguousarray(is_free, dtype=np.uint8)
        cost = None if cost is None else np.ascontiguousarray(cost, dtype=np.float64)
        first = np.empty(t, dtype=np.int32) if is_free is not None else None
        score = np.empty(t, dtype=np.float64) if cost is not None else None
    ctx._check(ctx._lib.cover_trajectory_reduce(ctx._h, C.c_void_p(_ptr(is_free)), C.c_void_p(_ptr(cost)), C.c_int32(_MEM_DEVICE if dev else _MEM_HOST),
                                                C.c_int64(t), C.c_int32(steps), C.c_void_p(_ptr(first)), C.c_void_p(_ptr(score))))
    return first, score


def get_radius(polygon) -> float:
    """get_radius(polygon) (src/cover/footprints.cpp:55-77): min distance from the centroid to the polygon's edges."""
    lib = _load()
    lib.cover_get_radius.restype = C.c_double
    xy = polygon_xy(polygon)
    return float(lib.cover_get_radius(C.c_void_p(xy.ctypes.data), C.c_int32(xy.size // 2)))


def split_polygon(polygon, radius: Optional[float] = None, slack: float = 0.0, max_vertices: int = 1 << 16, max_polygons: int = 256):
    """continuous_footprint(radius, polygon) without boost::geometry (footprints.cpp:79-128; one-argument form:
    radius = get_radius(polygon)): returns (outlines, areas), two lists of [2, V] polygons - the always-inscribed
    sub-regions whose outline cells are checked for 253/254 and the remainder that is scanned for 254 - ready for
    `SplitFootprint(ctx, outlines, areas)`.  Raises ValueError for a non-positive radius (std::invalid_argument upstream).
    `slack` (extension, metres; 0 = the reference's construction): the outlines eroded that much less, see cover_b200.h."""
    lib = _load()
    xy = polygon_xy(polygon)
    if radius is None:
        radius = get_radius(polygon)
    bufs = [np.empty(2 * max_vertices, dtype=np.float64) for _ in range(2)]
    starts = [np.zeros(max_polygons + 1, dtype=np.int32) for _ in range(2)]
    counts = [C.c_int32(0), C.c_int32(0)]
    rc = lib.cover_split_polygon_slack(C.c_double(radius), C.c_double(slack), C.c_void_p(xy.ctypes.data), C.c_int32(xy.size // 2),
                                 C.c_void_p(bufs[0].ctypes.data), C.c_void_p(starts[0].ctypes.data), C.byref(counts[0]),
                                 C.c_void_p(bufs[1].ctypes.data), C.c_void_p(starts[1].ctypes.data), C.byref(counts[1]),
                                 C.c_int32(max_vertices), C.c_int32(max_polygons))
    if rc == -1:
        raise ValueError("Radius must be positive")
    if rc != 0:
        raise CoverError(rc, "cover_split_polygon failed (more polygons / vertices than the buffers hold?)")
    out = []
    for b, st, n in zip(bufs, starts, counts):
        out.append([b[2 * st[k]:2 * st[k + 1]].reshape(-1, 2).T.copy() for k in range(n.value)])
    return out[0], out[1]


class Costmap:
    """Device copy of a costmap_2d::Costmap2D: uint8 [size_y, size_x] row-major + resolution/origin."""

    def __init__(self, ctx: Context, data, resolution: float, origin_x: float = 0.0, origin_y: float = 0.0):
        self.ctx = ctx
        shape = tuple(data.shape)
        if len(shape) != 2:
            raise ValueError("costmap data must be [size_y, size_x]")
        self.size_y, self.size_x = int(shape[0]), int(shape[1])
        self.resolution, self.origin_x, self.origin_y = float(resolution), float(origin_x), float(origin_y)
        h = C.c_void_p()
        ctx._check(ctx._lib.cover_map_create(ctx._h, C.c_int32(self.size_x), C.c_int32(self.size_y), C.c_double(resolution),
                                             C.c_double(origin_x), C.c_double(origin_y), C.byref(h)))
        self._h = h
        self.upload(data)

    def upload(self, data):
        """Full re-upload (host numpy array or device tensor)."""
        if _is_device(data):
            if tuple(data.shape) != (self.size_y, self.size_x) or not data.is_contiguous() or data.element_size() != 1:
                raise ValueError("device costmap must be a contiguous uint8 [size_y, size_x] tensor")
            self.ctx._check(self.ctx._lib.cover_map_upload(self._h, C.c_void_p(_ptr(data)), C.c_int32(_MEM_DEVICE)))
        else:
            arr = np.ascontiguousarray(data, dtype=np.uint8)
            if arr.shape != (self.size_y, self.size_x):
                raise ValueError("costmap shape mismatch")
            self.ctx._check(self.ctx._lib.cover_map_upload(self._h, C.c_void_p(arr.ctypes.data), C.c_int32(_MEM_HOST)))

    def upload_async(self, data: np.ndarray):
        """Queue a full re-upload from HOST memory and return at once (cover_map_upload_async): the copy runs as row
        bands on a second stream and later checks start as soon as the bands they need have landed.  `data` (uint8
        [size_y, size_x], C-contiguous; page-locked for a truly asynchronous copy) must stay alive and unchanged
        until the next call that returns results to the host, or `Context.synchronize()`."""
        if not isinstance(data, np.ndarray) or data.dtype != np.uint8 or not data.flags.c_contiguous or data.shape != (self.size_y, self.size_x):
            raise ValueError("upload_async needs a C-contiguous uint8 [size_y, size_x] numpy array (no conversion copy is made)")
        self._pending_upload = data  # keep the buffer alive while the copy is in flight
        self.ctx._check(self.ctx._lib.cover_map_upload_async(self._h, C.c_void_p(data.ctypes.data)))

    def upload_window(self, full_image: np.ndarray, x0: int, y0: int, w: int, h: int):
        arr = np.ascontiguousarray(full_image, dtype=np.uint8)
        if arr.shape != (self.size_y, self.size_x):
            raise ValueError("costmap shape mismatch")
        self.ctx._check(self.ctx._lib.cover_map_upload_window(self._h, C.c_void_p(arr.ctypes.data), C.c_int32(x0), C.c_int32(y0),
                                                              C.c_int32(w), C.c_int32(h)))

    def inflate(self, radius_cells: int, cost_by_d2: np.ndarray):
        """Inflate the lethal cells of the device map in place (see workloads.inflation_lut)."""
        lut = np.ascontiguousarray(cost_by_d2, dtype=np.uint8)
        if lut.size != radius_cells * radius_cells + 1:
            raise ValueError("cost_by_d2 must have radius_cells^2 + 1 entries")
        self.ctx._check(self.ctx._lib.cover_map_inflate(self._h, C.c_int32(radius_cells), C.c_void_p(lut.ctypes.data)))

    def download(self) -> np.ndarray:
        """Device map -> host uint8 [size_y, size_x] (e.g. after `inflate`)."""
        out = np.empty((self.size_y, self.size_x), dtype=np.uint8)
        self.ctx._check(self.ctx._lib.cover_map_download(self._h, C.c_void_p(out.ctypes.data)))
        return out

    def close(self):
        if getattr(self, "_h", None) and getattr(self.ctx, "_h", None):
            self.ctx._lib.cover_map_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- polygon x poses (path A) --------------------------------------------------------------
    def _polygon(self, fn: str, polygon, poses, first, count, out, want_cost, offset=None):
        ctx = self.ctx
        xy = polygon_xy(polygon)
        p, keep = ctx._poses(poses, first, count, offset)
        r, res = ctx._results(int(p.count), True, want_cost, True, out)
        ctx._check(getattr(ctx._lib, fn)(ctx._h, self._h, C.c_void_p(xy.ctypes.data), C.c_int32(xy.size // 2), C.byref(p), C.byref(r)))
        del keep
        return res

    def area_multi(self, outlines: Sequence, poses, op: str = "is_free", first=0, count=None, out=None) -> Result:
        """Batched is_free / accumulate_cost / max_cost over area_generator(res, polygon_vec): `outlines` is a list
        of [2, V_k] polygons, nested ones are holes (generators.cpp:126-135)."""
        ctx = self.ctx
        xy, starts = _pack(outlines)
        p, keep = ctx._poses(poses, first, count)
        code = {"is_free": 0, "accumulate_cost": 1, "max_cost": 2}[op]
        r, res = ctx._results(int(p.count), True, code != 0, True, out)
        ctx._check(ctx._lib.cover_area_multi_check(ctx._h, self._h, C.c_void_p(xy.ctypes.data), C.c_void_p(starts.ctypes.data), C.c_int32(len(outlines)),
                                                   C.c_int32(code), C.byref(p), C.byref(r)))
        del keep
        return res

    def area_is_free(self, polygon, poses, first=0, count=None, out=None, offset=None) -> Result:
        """Batched is_free(polygon, pose, map) (costmap.cpp:78-85); poses=None is is_free(polygon, map)."""
        return self._polygon("cover_area_is_free", polygon, poses, first, count, out, False, offset)

    def area_accumulate_cost(self, polygon, poses, first=0, count=None, out=None, offset=None) -> Result:
        """Batched accumulate_cost(area_generator(res, pose*polygon - origin), map)."""
        return self._polygon("cover_area_accumulate_cost", polygon, poses, first, count, out, True, offset)

    def area_max_cost(self, polygon, poses, first=0, count=None, out=None, offset=None) -> Result:
        return self._polygon("cover_area_max_cost", polygon, poses, first, count, out, True, offset)

    def outline_is_free(self, polygon, poses, first=0, count=None, out=None, offset=None) -> Result:
        """Batched is_free(outline_generator(res, pose*polygon - origin), map)."""
        return self._polygon("cover_outline_is_free", polygon, poses, first, count, out, False, offset)

    def outline_accumulate_cost(self, polygon, poses, first=0, count=None, out=None, offset=None) -> Result:
        return self._polygon("cover_outline_accumulate_cost", polygon, poses, first, count, out, True, offset)

    def outline_max_cost(self, polygon, poses, first=0, count=None, out=None, offset=None) -> Result:
        return self._polygon("cover_outline_max_cost", polygon, poses, first, count, out, True, offset)

    # -- expand (to_outline / to_area, batched) ------------------------------------------------
    def _expand(self, fn: str, polygon, poses, first, count):
        ctx = self.ctx
        xy = polygon_xy(polygon)
        p, keep = ctx._poses(poses, first, count)
        n = int(p.count)
        starts = np.zeros(n + 1, dtype=np.int64)
        status = np.zeros(max(n, 1), dtype=np.uint8)
        call = getattr(ctx._lib, fn)
        args = (ctx._h, self._h, C.c_void_p(xy.ctypes.data), C.c_int32(xy.size // 2), C.byref(p))
        ctx._check(call(*args, None, C.c_int64(0), C.c_void_p(starts.ctypes.data), C.c_void_p(status.ctypes.data), C.c_int32(_MEM_HOST)))
        total = int(starts[n])
        cells = np.empty((max(total, 1), 2), dtype=np.int32)
        ctx._check(call(*args, C.c_void_p(cells.ctypes.data), C.c_int64(total), C.c_void_p(starts.ctypes.data),
                        C.c_void_p(status.ctypes.data), C.c_int32(_MEM_HOST)))
        del keep
        return cells[:total], starts, status[:n]

    def area_cells(self, polygon, poses, first=0, count=None):
        """Batched to_area(res, pose*polygon - origin) (expand.hpp:75-80): (cells [total, 2] in generator order,
        starts [N + 1], status [N])."""
        return self._expand("cover_area_expand", polygon, poses, first, count)

    def outline_cells(self, polygon, poses, first=0, count=None):
        """Batched to_outline(res, pose*polygon - origin) (expand.hpp:68-72)."""
        return self._expand("cover_outline_expand", polygon, poses, first, count)

    # -- rays ----------------------------------------------------------------------------------
    def _rays(self, fn: str, segments, offset, out, want_cost):
        ctx = self.ctx
        if _is_device(segments):
            seg, n, mem = segments, int(segments.shape[0]), _MEM_DEVICE
        else:
            seg = np.ascontiguousarray(segments, dtype=np.int32).reshape(-1, 4)
            n, mem = len(seg), _MEM_HOST
        off = None if offset is None else np.ascontiguousarray(offset, dtype=np.int32).reshape(2)
        r, res = ctx._results(n, True, want_cost, True, out)
        ctx._check(getattr(ctx._lib, fn)(ctx._h, self._h, C.c_void_p(_ptr(seg)), C.c_int32(mem), C.c_int64(n),
                                         C.c_void_p(_ptr(off)), C.byref(r)))
        return res

    def ray_is_free(self, segments, offset=None, out=None) -> Result:
        """Batched is_free(ray_generator(begin, end)[, offset], map); segments [N, 4] = x0,y0,x1,y1 (end exclusive)."""
        return self._rays("cover_ray_is_free", segments, offset, out, False)

    def ray_accumulate_cost(self, segments, offset=None, out=None) -> Result:
        return self._rays("cover_ray_accumulate_cost", segments, offset, out, True)

    def ray_max_cost(self, segments, offset=None, out=None) -> Result:
        return self._rays("cover_ray_max_cost", segments, offset, out, True)


def _offsets(offsets):
    if _is_device(offsets):
        return offsets, int(offsets.shape[0]), _MEM_DEVICE
    arr = np.ascontiguousarray(offsets, dtype=np.int32).reshape(-1, 2)
    return arr, len(arr), _MEM_HOST


class DiscretePolygon:
    """A cached cover::discrete_polygon on the device (is_free(dp, offset, map), costmap.cpp:64-69)."""

    def __init__(self, ctx: Context, cells):
        self.ctx = ctx
        arr = np.ascontiguousarray(cells, dtype=np.int32).reshape(-1, 2)
        h = C.c_void_p()
        ctx._check(ctx._lib.cover_cells_create(ctx._h, C.c_void_p(arr.ctypes.data), C.c_int64(len(arr)), C.byref(h)))
        self._h = h
        self.n_cells = len(arr)

    def _run(self, fn, cmap: Costmap, offsets, out, want_cost):
        ctx = self.ctx
        off, n, mem = _offsets(offsets)
        r, res = ctx._results(n, True, want_cost, True, out)
        ctx._check(getattr(ctx._lib, fn)(ctx._h, cmap._h, self._h, C.c_void_p(_ptr(off)), C.c_int32(mem), C.c_int64(n), C.byref(r)))
        return res

    def is_free(self, cmap: Costmap, offsets, out=None) -> Result:
        return self._run("cover_cells_is_free", cmap, offsets, out, False)

    def accumulate_cost(self, cmap: Costmap, offsets, out=None) -> Result:
        return self._run("cover_cells_accumulate_cost", cmap, offsets, out, True)

    def max_cost(self, cmap: Costmap, offsets, out=None) -> Result:
        return self._run("cover_cells_max_cost", cmap, offsets, out, True)

    def close(self):
        if getattr(self, "_h", None) and getattr(self.ctx, "_h", None):
            self.ctx._lib.cover_cells_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class FootprintBank:
    """K cover::discrete_footprint's (one per heading bin) on the device; `is_free(offsets, which)` is the
    batched discrete_footprint::is_free(offset, map) (footprints.cpp:240-247)."""

    def __init__(self, ctx: Context, footprints: Sequence):
        """footprints: sequence of (outline_cells [No,2], area_cells [Na,2])."""
        self.ctx = ctx
        k = len(footprints)
        o_starts = np.zeros(k + 1, dtype=np.int64)
        a_starts = np.zeros(k + 1, dtype=np.int64)
        oc, ac = [], []
        for i, (o, a) in enumerate(footprints):
            o = np.asarray(o, dtype=np.int32).reshape(-1, 2)
            a = np.asarray(a, dtype=np.int32).reshape(-1, 2)
            oc.append(o)
            ac.append(a)
            o_starts[i + 1] = o_starts[i] + len(o)
            a_starts[i + 1] = a_starts[i] + len(a)
        ocat = np.ascontiguousarray(np.concatenate(oc)) if k else np.zeros((0, 2), np.int32)
        acat = np.ascontiguousarray(np.concatenate(ac)) if k else np.zeros((0, 2), np.int32)
        h = C.c_void_p()
        ctx._check(ctx._lib.cover_footprints_create(ctx._h, C.c_void_p(ocat.ctypes.data), C.c_void_p(o_starts.ctypes.data),
                                                    C.c_void_p(acat.ctypes.data), C.c_void_p(a_starts.ctypes.data), C.c_int32(k),
                                                    C.byref(h)))
        self._h = h
        self.n_footprints = k

    def is_free(self, cmap: Costmap, offsets, which=None, out=None) -> Result:
        ctx = self.ctx
        off, n, mem = _offsets(offsets)
        if which is not None and not _is_device(which):
            which = np.ascontiguousarray(which, dtype=np.int32)
        r, res = ctx._results(n, True, False, True, out)
        ctx._check(ctx._lib.cover_footprints_is_free(ctx._h, cmap._h, self._h, C.c_void_p(_ptr(off)), C.c_void_p(_ptr(which)),
                                                     C.c_int32(mem), C.c_int64(n), C.byref(r)))
        return res

    def close(self):
        if getattr(self, "_h", None) and getattr(self.ctx, "_h", None):
            self.ctx._lib.cover_footprints_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _pack(polys):
    starts = np.zeros(len(polys) + 1, dtype=np.int32)
    flat = []
    for k, p in enumerate(polys):
        xy = polygon_xy(p)
        flat.append(xy)
        starts[k + 1] = starts[k] + xy.size // 2
    xy = np.concatenate(flat) if flat else np.zeros(0)
    return np.ascontiguousarray(xy, dtype=np.float64), starts


class SplitFootprint:
    """cover::continuous_footprint given as (outlines, areas) (footprints.hpp:136 public constructor);
    `is_free(cmap, poses)` is the batched continuous_footprint::is_free(pose, map) (footprints.cpp:130-196)."""

    def __init__(self, ctx: Context, outlines: Sequence, areas: Sequence):
        self.ctx = ctx
        oxy, ost = _pack(outlines)
        axy, ast = _pack(areas)
        h = C.c_void_p()
        ctx._check(ctx._lib.cover_split_create(ctx._h, C.c_void_p(oxy.ctypes.data), C.c_void_p(ost.ctypes.data), C.c_int32(len(outlines)),
                                               C.c_void_p(axy.ctypes.data), C.c_void_p(ast.ctypes.data), C.c_int32(len(areas)), C.byref(h)))
        self._h = h

    def is_free(self, cmap: Costmap, poses, first=0, count=None, out=None) -> Result:
        ctx = self.ctx
        p, keep = ctx._poses(poses, first, count)
        r, res = ctx._results(int(p.count), True, False, True, out)
        ctx._check(ctx._lib.cover_split_is_free(ctx._h, cmap._h, self._h, C.byref(p), C.byref(r)))
        del keep
        return res

    def close(self):
        if getattr(self, "_h", None) and getattr(self.ctx, "_h", None):
            self.ctx._lib.cover_split_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
