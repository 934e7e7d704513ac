"""GPU parity at the FULL sizes of BASELINE.json's configs, against the reference's own code (oracle/_ref) when it was
built, else the restatement: C2 (2000^2 inflated map, 1e6 poses: outlines + rays), C3 (4000^2, 72-heading potato bank
x 1e7 offsets + the split footprint's pose form), C4 (all eight lattice windows bench.py hands its ranks, >= 2 M sampled
poses each, the axis-aligned headings included), C5 (8000^2, all eight shapes).  The CPU side is sampled where the full
batch would take minutes; the GPU side always runs the whole batch.
"""
import math

import numpy as np
import pytest

from oracle import pyoracle as po
from tests.fixtures import FOOTPRINTS

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cb():
    import cover_b200
    return cover_b200


@pytest.fixture(scope="module")
def ctx(cb):
    c = cb.Context(0)
    yield c
    c.close()


def cpu():
    ref = po.reference()
    return ref if ref is not None else po.default()


def same(gpu_free, gpu_status, r, what):
    """Verdicts wherever both sides finished + the std::logic_error flag (the reference knows no other status)."""
    gpu_free, gpu_status = np.asarray(gpu_free), np.asarray(gpu_status)
    np.testing.assert_array_equal(gpu_status == 1, r.status == 1, err_msg=f"logic_error flags: {what}")
    ok = (gpu_status == 0) & (r.status == 0)
    np.testing.assert_array_equal(gpu_free[ok], r.is_free[ok], err_msg=what)
    assert ok.mean() > 0.99, what


@pytest.fixture(scope="module")
def map4000():
    from cover_b200 import workloads as wl
    return wl.inflated_costmap(4000, 4000, 0.05, seed=3)


def test_c2_outlines_and_rays_full_size(cb, ctx):
    from cover_b200 import workloads as wl
    n = 1_000_000
    data = wl.inflated_costmap(2000, 2000, 0.05, seed=3)
    cmap = cb.Costmap(ctx, data, 0.05)
    omap = po.Costmap(data, 0.05)
    poses = wl.random_poses(n, seed=4, lo=0.2, hi=99.8)  # some poses reach over the map border
    for name, poly in (("rect", wl.RECT), ("TORU", FOOTPRINTS[5])):
        for op_name, op in (("is_free", po.OP_IS_FREE), ("accumulate_cost", po.OP_ACCUMULATE)):
            gpu = getattr(cmap, f"outline_{op_name}")(poly, poses)
            r = cpu().polygon_batch(omap, poly, poses, shape=po.SHAPE_OUTLINE, op=op, threads=16)
            same(gpu.is_free, gpu.status, r, f"C2 {name} outline {op_name}")
            if op == po.OP_ACCUMULATE:
                np.testing.assert_array_equal(gpu.cost, r.cost)
        gpu = cmap.area_is_free(poly, poses)
        same(gpu.is_free, gpu.status, cpu().polygon_batch(omap, poly, poses, threads=16), f"C2 {name} area")
    rng = np.random.default_rng(5)
    b = rng.integers(-10, 2010, (n, 2))
    segs = np.concatenate([b, b + rng.integers(-25, 26, (n, 2))], 1).astype(np.int32)
    for op_name, op in (("is_free", po.OP_IS_FREE), ("accumulate_cost", po.OP_ACCUMULATE)):
        gpu = getattr(cmap, f"ray_{op_name}")(segs)
        r = cpu().rays_batch(omap, segs, op=op, threads=16)
        np.testing.assert_array_equal(gpu.is_free, r.is_free)
        if op == po.OP_ACCUMULATE:
            np.testing.assert_array_equal(gpu.cost, r.cost)


def test_c3_footprint_bank_and_split_full_size(cb, ctx, map4000):
    from cover_b200 import workloads as wl
    n = 10_000_000
    cmap = cb.Costmap(ctx, map4000, 0.05)
    omap = po.Costmap(map4000, 0.05)
    pot = wl.potato(48)
    outlines, areas = [wl.scaled(pot, 0.45)], [pot]
    bank = [cpu().make_discrete_footprint(0.05, [wl.rotation(-math.pi + 2 * math.pi * k / 72) @ o for o in outlines],
                                          [wl.rotation(-math.pi + 2 * math.pi * k / 72) @ a for a in areas]) for k in range(72)]
    fb = cb.FootprintBank(ctx, bank)
    rng = np.random.default_rng(5)
    offs = rng.integers(-20, 4020, (n, 2)).astype(np.int32)
    which = rng.integers(0, 72, n).astype(np.int32)
    gpu = fb.is_free(cmap, offs, which)
    m = 1_500_000  # >= 1e6 compared: the head, the tail and a random sample in between
    idx = np.concatenate([np.arange(500_000), np.arange(n - 500_000, n), rng.choice(n, 500_000, replace=False)])
    assert idx.size == m
    r = cpu().discrete_footprint_batch(omap, bank, offs[idx], which[idx], threads=16)
    np.testing.assert_array_equal(gpu.is_free[idx], r.is_free)
    assert 0.05 < r.is_free.mean() < 0.95
    # the split footprint under arbitrary poses (continuous_footprint::is_free, footprints.cpp:130-196)
    sf = cb.SplitFootprint(ctx, outlines, areas)
    poses = wl.random_poses(1_000_000, seed=6, lo=0.5, hi=199.5)
    gpu = sf.is_free(cmap, poses)
    r = cpu().continuous_footprint_batch(omap, outlines, areas, poses, threads=16)
    np.testing.assert_array_equal(gpu.is_free, r.is_free)


def test_c4_all_rank_windows(cb, ctx, map4000):
    """Every window bench.py gives a rank (lattice_for_rank): the whole 100 252 800-pose lattice on the GPU, >= 2 M poses
    of it on the CPU - 250 000-pose chunks at random offsets plus, for each of the four axis-aligned headings, a chunk from
    the middle of its plane."""
    from bench import lattice_for_rank
    from cover_b200 import workloads as wl
    cmap = cb.Costmap(ctx, map4000, 0.05)
    omap = po.Costmap(map4000, 0.05)
    for rank in range(8):
        desc = lattice_for_rank(rank)
        lat, olat = cb.Lattice(**desc), po.Lattice(**desc)
        n = lat.count
        out = cb.Result(free_bits=np.zeros((n + 31) // 32, dtype=np.uint32), not_ok=np.zeros(1, dtype=np.int64))
        cmap.area_is_free(wl.RECT, lat, out=out)
        free = out.unpack_bits(n)
        assert int(out.not_ok[0]) == 0
        rng = np.random.default_rng(1000 + rank)
        chunk = 250_000
        per_heading = desc["nx"] * desc["ny"]
        starts = rng.integers(0, n - chunk, 5).tolist() + [h * per_heading + per_heading // 2 for h in (0, 18, 36, 54)]  # -pi, -pi/2, 0, pi/2
        total = 0
        for s in starts:
            r = cpu().polygon_batch(omap, wl.RECT, olat, first=int(s), count=chunk, threads=16)
            np.testing.assert_array_equal(free[s:s + chunk], r.is_free, err_msg=f"window {rank}, poses {s}..")
            assert not r.status.any()
            total += chunk
        assert total >= 2_000_000
        # byte outputs of the same sweep agree with the bits
        full = cmap.area_is_free(wl.RECT, lat, first=per_heading * 17, count=3 * per_heading)
        np.testing.assert_array_equal(full.is_free, free[per_heading * 17:per_heading * 20])
        assert not full.status.any()


def test_c5_eight_shapes_full_size(cb, ctx):
    from cover_b200 import workloads as wl
    data = wl.inflated_costmap(8000, 8000, 0.05, seed=7)
    cmap = cb.Costmap(ctx, data, 0.05)
    omap = po.Costmap(data, 0.05)
    n = 1_000_000
    poses = wl.random_poses(n, seed=8, lo=1.0, hi=399.0)
    shapes = [("rect(4)", wl.RECT), ("L(6)", FOOTPRINTS[6]), ("S(6)", FOOTPRINTS[0]), ("W(8)", FOOTPRINTS[1]), ("M(10)", FOOTPRINTS[2]),
              ("SOTO(22)", FOOTPRINTS[4]), ("TORU(32)", FOOTPRINTS[5]), ("potato(64)", wl.potato(64))]
    m = 100_000
    for name, poly in shapes:
        gpu = cmap.area_is_free(poly, poses)
        r = cpu().polygon_batch(omap, poly, poses[:m], threads=16)
        same(gpu.is_free[:m], gpu.status[:m], r, f"C5 {name}")
        assert 0.02 < r.is_free.mean() < 0.999, name
    # the same shapes on a lattice window of the 8000^2 map (the bit-parallel kernels: any polygon, any height)
    desc = wl.dwa_lattice(nx=400, ny=320, n_theta=12, x0=150.0, y0=200.0)
    for name, poly in shapes:
        gpu = cmap.area_is_free(poly, cb.Lattice(**desc))
        first, count = 400 * 320 * 5 + 777, 120_000
        r = cpu().polygon_batch(omap, poly, po.Lattice(**desc), first=first, count=count, threads=16)
        same(gpu.is_free[first:first + count], gpu.status[first:first + count], r, f"C5 lattice {name}")
