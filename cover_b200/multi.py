"""Pose sharding across the GPUs of one box, as a library feature.

Poses are independent units: every device gets a contiguous block of the pose range (`sharding.partition`) and a
replica of the costmap, one host thread drives each device's context (ctypes releases the GIL for the duration of a
call), and every block's results land in ITS slice of one host array.  No collective, no exchange between devices.
The same device may be listed several times: the shards then run as independent contexts (= streams) of that device,
which is also how the sharding logic is tested on a one-GPU box.  C++ mirror: `cover_b200::multi_gpu`
(include/cover_b200.hpp).
"""
from __future__ import annotations

import ctypes as C
import threading
from typing import List, Optional, Sequence

import numpy as np

from . import Context, Costmap, Lattice, Result, _load
from .sharding import partition


def device_count() -> int:
    lib = _load()
    lib.cover_device_count.restype = C.c_int
    return int(lib.cover_device_count())


def _each(n: int, body):
    """body(k) for k in range(n), one host thread each; the first exception is re-raised here."""
    errors: List[Optional[BaseException]] = [None] * n

    def run(k):
        try:
            body(k)
        except BaseException as e:  # noqa: BLE001 - re-raised below
            errors[k] = e
    threads = [threading.Thread(target=run, args=(k,)) for k in range(n)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for e in errors:
        if e is not None:
            raise e


class MultiGPU:
    """One Context per listed device (default: every device the process sees)."""

    def __init__(self, devices: Optional[Sequence[int]] = None):
        devices = list(range(device_count())) if devices is None else list(devices)
        if not devices:
            raise RuntimeError("cover_b200: no usable CUDA device (there is no CPU fallback)")
        self.devices = devices
        self.contexts = [Context(d) for d in devices]

    def __len__(self) -> int:
        return len(self.contexts)

    def costmap(self, data: np.ndarray, resolution: float, origin_x: float = 0.0, origin_y: float = 0.0) -> "ShardedCostmap":
        return ShardedCostmap(self, data, resolution, origin_x, origin_y)

    def close(self):
        for c in self.contexts:
            c.close()


class ShardedCostmap:
    """The costmap replicated on every device of a MultiGPU; checks split their poses across the replicas."""

    def __init__(self, gpus: MultiGPU, data: np.ndarray, resolution: float, origin_x: float, origin_y: float):
        self.gpus = gpus
        data = np.ascontiguousarray(data, dtype=np.uint8)
        self.maps: List[Optional[Costmap]] = [None] * len(gpus)

        def make(k):
            gpus.contexts[k].bind()  # (the shard's thread stays on its device: no switch from device 0 and back per call)
            self.maps[k] = Costmap(gpus.contexts[k], data, resolution, origin_x, origin_y)
        _each(len(gpus), make)

    def upload(self, data: np.ndarray):
        data = np.ascontiguousarray(data, dtype=np.uint8)

        def run(k):
            self.gpus.contexts[k].bind()
            self.maps[k].upload(data)
        _each(len(self.maps), run)

    def _blocks(self, count: int, granule: int = 1):
        """(first, n) per shard; with granule > 1 the borders are multiples of it (whole words of packed verdicts)."""
        units = -(-count // granule)
        out = []
        for k in range(len(self.maps)):
            f, n = partition(units, len(self.maps), k)
            first = min(f * granule, count)
            out.append((first, max(0, min(n * granule, count - first))))
        return out

    def _polygon(self, name: str, polygon, poses, want_cost: bool) -> Result:
        lattice = isinstance(poses, Lattice)
        if not lattice:
            poses = np.ascontiguousarray(poses, dtype=np.float64).reshape(-1, 4)
        count = poses.count if lattice else len(poses)
        res = Result(np.empty(count, dtype=np.uint8), np.empty(count, dtype=np.float64) if want_cost else None, np.empty(count, dtype=np.uint8))
        blocks = self._blocks(count)

        def run(k):
            first, n = blocks[k]
            if n == 0:
                return
            self.gpus.contexts[k].bind()
            out = Result(res.is_free[first:first + n], None if res.cost is None else res.cost[first:first + n], res.status[first:first + n])
            fn = getattr(self.maps[k], name)
            if lattice:
                fn(polygon, poses, first=first, count=n, out=out)
            else:
                fn(polygon, poses[first:first + n], out=out)
        _each(len(self.maps), run)
        return res

    def area_is_free(self, polygon, poses) -> Result:
        """Batched is_free(polygon, pose, map) (costmap.cpp:78-85), the poses split across the devices."""
        return self._polygon("area_is_free", polygon, poses, False)

    def area_accumulate_cost(self, polygon, poses) -> Result:
        return self._polygon("area_accumulate_cost", polygon, poses, True)

    def outline_is_free(self, polygon, poses) -> Result:
        return self._polygon("outline_is_free", polygon, poses, False)

    def area_is_free_bits(self, polygon, lattice: Lattice) -> Result:
        """Lattice sweep with bit-packed verdicts: every shard writes whole words of ONE packed array."""
        count = lattice.count
        bits = np.zeros((count + 31) // 32, dtype=np.uint32)
        bad = np.zeros(len(self.maps), dtype=np.int64)
        blocks = self._blocks(count, 32)

        def run(k):
            first, n = blocks[k]
            if n:
                self.gpus.contexts[k].bind()
                self.maps[k].area_is_free(polygon, lattice, first=first, count=n, out=Result(free_bits=bits[first // 32:], not_ok=bad[k:k + 1]))
        _each(len(self.maps), run)
        return Result(free_bits=bits, not_ok=np.array([bad.sum()], dtype=np.int64))
