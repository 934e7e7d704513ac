"""The N > 1 path on CPU: world size 2, gloo backend.  Every rank checks its contiguous block of the pose
range (here with the CPU oracle standing in for the per-rank GPU — this test is about the host-side
sharding / gather logic), the slices are gathered, and the result must equal the unsharded run."""
import math
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cover_b200 import sharding
from cover_b200 import workloads as wl


def test_partition_is_a_disjoint_cover():
    for count in (0, 1, 7, 100, 100_252_800):
        for world in (1, 2, 3, 4, 8):
            blocks = [sharding.partition(count, world, r) for r in range(world)]
            assert blocks[0][0] == 0
            for (f0, n0), (f1, _) in zip(blocks, blocks[1:]):
                assert f0 + n0 == f1
            assert blocks[-1][0] + blocks[-1][1] == count
            assert max(n for _, n in blocks) - min(n for _, n in blocks) <= 1
            assert max(n for _, n in blocks) <= sharding.max_block(count, world) or count == 0
    with pytest.raises(ValueError):
        sharding.partition(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    from oracle import pyoracle as po
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        data = wl.inflated_costmap(300, 300, 0.05, seed=3, obstacle_p=1e-3)  # replicated on every rank
        th = -math.pi + 2 * math.pi * np.arange(5) / 5
        lat = po.Lattice(1.0, 0.05, 101, 1.5, 0.05, 77, np.stack([np.cos(th), np.sin(th)], 1))
        omap = po.Costmap(data, 0.05)
        first, n = sharding.partition(lat.count, world, rank)
        local = po.polygon_batch(omap, wl.RECT, lat, first=first, count=n)       # lattice: first/count
        full = sharding.gather_slices(torch.from_numpy(local.is_free), lat.count)
        poses = wl.random_poses(10_001, seed=9, lo=0.5, hi=14.5)
        pf, pn = sharding.partition(len(poses), world, rank)
        local2 = po.polygon_batch(omap, wl.potato(), poses[pf:pf + pn], op=po.OP_ACCUMULATE)  # explicit poses: slices
        full2 = sharding.gather_slices(torch.from_numpy(local2.cost), len(poses))
        if rank == 0:
            whole = po.polygon_batch(omap, wl.RECT, lat)
            whole2 = po.polygon_batch(omap, wl.potato(), poses, op=po.OP_ACCUMULATE)
            ret["ok"] = bool(np.array_equal(full.numpy(), whole.is_free) and np.array_equal(full2.numpy(), whole2.cost, equal_nan=True))
            ret["free"] = float(whole.is_free.mean())
    finally:
        dist.destroy_process_group()


def test_sharded_checks_equal_unsharded_world2():
    ctx = mp.get_context("spawn")
    with ctx.Manager() as manager:
        ret = manager.dict()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(timeout=240)
            assert p.exitcode == 0
        assert ret.get("ok") is True
        assert 0.0 < ret["free"] < 1.0
