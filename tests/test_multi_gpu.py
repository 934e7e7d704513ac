"""Pose sharding as a library feature (cover_b200.multi.MultiGPU, C++ mirror cover_b200::multi_gpu): one context per
device, costmap replicated, contiguous pose blocks, results gathered into one host array - against the single-GPU
result and an oracle sample.  On a box with one GPU the shards are three contexts of that device (the partition /
gather logic is the same); with `gpurun --gpus N` they are N real devices."""
import os
import subprocess

import numpy as np
import pytest

from oracle import pyoracle as po
from tests.fixtures import FOOTPRINTS, RECT

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def devices():
    import torch
    n = torch.cuda.device_count()
    return list(range(n)) if n >= 2 else [0, 0, 0]


def test_sharded_lattice_and_c5_sweep_match_single_gpu_and_oracle():
    import cover_b200 as cb
    from cover_b200 import workloads as wl
    from cover_b200.multi import MultiGPU, device_count
    assert device_count() >= 1
    data = wl.inflated_costmap(2000, 2000, 0.05, seed=21, obstacle_p=4e-4)
    gpus = MultiGPU(devices())
    smap = gpus.costmap(data, 0.05)
    single_ctx = cb.Context(0)
    single = cb.Costmap(single_ctx, data, 0.05)
    omap = po.Costmap(data, 0.05)
    cpu = po.reference() or po.default()
    # a lattice whose block borders fall inside rows and inside headings
    desc = wl.dwa_lattice(nx=701, ny=333, n_theta=7, x0=10.0, y0=12.0)
    lat = cb.Lattice(**desc)
    want = single.area_is_free(RECT, lat)
    got = smap.area_is_free(RECT, lat)
    np.testing.assert_array_equal(got.is_free, want.is_free)
    np.testing.assert_array_equal(got.status, want.status)
    bits = smap.area_is_free_bits(RECT, lat)
    np.testing.assert_array_equal(bits.unpack_bits(lat.count), want.is_free)
    assert int(bits.not_ok[0]) == 0
    first, count = 701 * 333 * 3 + 12345, 200_000
    r = cpu.polygon_batch(omap, RECT, po.Lattice(**desc), first=first, count=count, threads=8)
    np.testing.assert_array_equal(got.is_free[first:first + count], r.is_free)
    # the C5 sweep: eight shapes x random poses, explicit pose arrays split by pointer offset
    poses = wl.random_poses(300_001, seed=22, lo=0.5, hi=99.5)
    shapes = [RECT, FOOTPRINTS[6], FOOTPRINTS[0], FOOTPRINTS[1], FOOTPRINTS[2], FOOTPRINTS[4], FOOTPRINTS[5], wl.potato(64)]
    for poly in shapes:
        want = single.area_is_free(poly, poses)
        got = smap.area_is_free(poly, poses)
        np.testing.assert_array_equal(got.is_free, want.is_free)
        np.testing.assert_array_equal(got.status, want.status)
        r = cpu.polygon_batch(omap, poly, poses[-20_000:], threads=8)
        ok = got.status[-20_000:] == 0
        np.testing.assert_array_equal(got.is_free[-20_000:][ok], r.is_free[ok])
    acc = smap.area_accumulate_cost(RECT, poses)
    np.testing.assert_array_equal(acc.cost, single.area_accumulate_cost(RECT, poses).cost)
    # a new image reaches every replica
    data2 = wl.inflated_costmap(2000, 2000, 0.05, seed=23, obstacle_p=1e-3)
    smap.upload(data2)
    single.upload(data2)
    np.testing.assert_array_equal(smap.area_is_free(RECT, lat).is_free, single.area_is_free(RECT, lat).is_free)
    gpus.close()
    single_ctx.close()


def test_cpp_multi_gpu_helper(tmp_path):
    """tests/cpp/multi_gpu_test.cpp: cover_b200::multi_gpu against a single context, through the C++ host layer."""
    import cover_b200 as cb
    exe = str(tmp_path / "multi_gpu_test")
    lib_dir = os.path.dirname(cb.library_path())
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-pthread", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "multi_gpu_test.cpp"),
                    "-o", exe, "-L", lib_dir, "-lcover_b200", f"-Wl,-rpath,{lib_dir}"], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    assert "multi_gpu ok" in out, out
