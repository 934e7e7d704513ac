"""GPU parity: every C-ABI entry (through the ctypes host layer) against the CPU oracle on identical
seeded inputs.  Integer/byte work -> the bar is BIT-EXACT verdicts, costs and statuses.
Run on the B200 box: `python -m pytest tests -m gpu`.
"""
import math

import numpy as np
import pytest

from oracle import pyoracle as po
from tests.fixtures import FOOTPRINTS, FOOTPRINT_NAMES, RECT, head_polygons

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


def mixed_map(rng, sx, sy, lethal=0.002):
    data = rng.integers(0, 253, (sy, sx), dtype=np.uint8)
    r = rng.random((sy, sx))
    data[r < lethal] = 254
    data[(r >= lethal) & (r < 1.5 * lethal)] = 255
    data[(r >= 1.5 * lethal) & (r < 2.5 * lethal)] = 253
    return data


def poses_around(rng, n, lo, hi):
    th = rng.uniform(-math.pi, math.pi, n)
    return np.stack([np.cos(th), np.sin(th), rng.uniform(lo, hi, n), rng.uniform(lo, hi, n)], axis=1)


def assert_same(gpu, cpu, cost=False, what=""):
    np.testing.assert_array_equal(gpu.status, cpu.status, err_msg=f"status {what}")
    np.testing.assert_array_equal(gpu.is_free, cpu.is_free, err_msg=f"is_free {what}")
    if cost:
        np.testing.assert_array_equal(gpu.cost, cpu.cost, err_msg=f"cost {what}")  # NaN == NaN positionally


OPS = [("is_free", po.OP_IS_FREE, False), ("accumulate_cost", po.OP_ACCUMULATE, True), ("max_cost", po.OP_MAX, True)]


@pytest.mark.parametrize("name,op,has_cost", OPS)
def test_rect_area_random_poses(cb, ctx, name, op, has_cost):
    """C1 shape: 1.0 x 0.6 m rectangle, 5 cm, 1000 x 1000 map, random poses (some leave the map)."""
    rng = np.random.default_rng(1)
    data = mixed_map(rng, 1000, 1000)
    poses = poses_around(rng, 100_000, -0.5, 50.5)
    cmap = cb.Costmap(ctx, data, 0.05)
    gpu = getattr(cmap, f"area_{name}")(RECT, poses)
    cpu = po.polygon_batch(po.Costmap(data, 0.05), RECT, poses, shape=po.SHAPE_AREA, op=op, threads=8)
    assert_same(gpu, cpu, has_cost, name)
    assert 0.05 < cpu.is_free.mean() < 0.95


@pytest.mark.parametrize("name,op,has_cost", OPS)
def test_rect_outline_random_poses(cb, ctx, name, op, has_cost):
    rng = np.random.default_rng(2)
    data = mixed_map(rng, 700, 500, lethal=0.01)
    poses = poses_around(rng, 50_000, -0.5, 30.0)
    cmap = cb.Costmap(ctx, data, 0.05, -1.25, 0.7)
    gpu = getattr(cmap, f"outline_{name}")(RECT, poses)
    cpu = po.polygon_batch(po.Costmap(data, 0.05, -1.25, 0.7), RECT, poses, shape=po.SHAPE_OUTLINE, op=op, threads=8)
    assert_same(gpu, cpu, has_cost, name)


def test_rect_lattice_matches_oracle_and_explicit_poses(cb, ctx):
    """C4 shape at reduced size: lattice descriptor == the same poses materialised on the host == oracle."""
    rng = np.random.default_rng(3)
    data = mixed_map(rng, 400, 400, lethal=0.001)
    th = -math.pi + 2 * math.pi * np.arange(9) / 9
    cs = np.stack([np.cos(th), np.sin(th)], 1)
    lat = cb.Lattice(0.31, 0.05, 150, 0.47, 0.05, 120, cs)
    olat = po.Lattice(0.31, 0.05, 150, 0.47, 0.05, 120, cs)
    cmap = cb.Costmap(ctx, data, 0.05)
    gpu = cmap.area_is_free(RECT, lat)
    cpu = po.polygon_batch(po.Costmap(data, 0.05), RECT, olat, threads=8)
    assert_same(gpu, cpu)
    explicit = cmap.area_is_free(RECT, olat.poses_cs())
    np.testing.assert_array_equal(gpu.is_free, explicit.is_free)
    # a sub-range (pose sharding across GPUs uses first/count)
    part = cmap.area_is_free(RECT, lat, first=12_345, count=50_000)
    np.testing.assert_array_equal(part.is_free, gpu.is_free[12_345:62_345])
    acc = cmap.area_accumulate_cost(RECT, lat, first=1000, count=20_000)
    cpu_acc = po.polygon_batch(po.Costmap(data, 0.05), RECT, olat, op=po.OP_ACCUMULATE, first=1000, count=20_000, threads=8)
    assert_same(acc, cpu_acc, True)


@pytest.mark.parametrize("pitch", [0.05, 0.025, 0.1, 0.07])
def test_lattice_axis_aligned_headings_and_map_border(cb, ctx, pitch):
    """Headings where the vertices sit exactly on cell boundaries (the relative shape flips with rounding
    noise from pose to pose) on a lattice that sticks out of the map on every side, at cell pitch and at
    pitches that are not the cell size."""
    rng = np.random.default_rng(33)
    data = mixed_map(rng, 300, 240, lethal=0.002)
    th = np.array([0.0, math.pi / 2, math.pi, -math.pi / 2, 0.3, math.atan2(3, 4), 1e-9])
    cs = np.stack([np.cos(th), np.sin(th)], 1)
    nx, ny = int(17.0 / pitch), int(14.0 / pitch)
    nx, ny = min(nx, 400), min(ny, 330)
    for origin in ((0.0, 0.0), (-1.3, 0.65)):
        desc = dict(x0=origin[0] - 1.0, dx=pitch, nx=nx, y0=origin[1] - 1.0, dy=pitch, ny=ny, cs=cs)
        cmap = cb.Costmap(ctx, data, 0.05, *origin)
        gpu = cmap.area_is_free(RECT, cb.Lattice(**desc))
        cpu = po.polygon_batch(po.Costmap(data, 0.05, *origin), RECT, po.Lattice(**desc), threads=8)
        assert_same(gpu, cpu, what=f"pitch {pitch} origin {origin}")
        assert 0.05 < cpu.is_free.mean() < 0.95
        for name, op in (("accumulate_cost", po.OP_ACCUMULATE), ("max_cost", po.OP_MAX)):  # the value folds of the lattice kernel
            gpu = getattr(cmap, f"area_{name}")(RECT, cb.Lattice(**desc))
            cpu = po.polygon_batch(po.Costmap(data, 0.05, *origin), RECT, po.Lattice(**desc), op=op, threads=8)
            assert_same(gpu, cpu, True, what=f"{name} pitch {pitch} origin {origin}")


@pytest.mark.parametrize("geom", [dict(nx=70, ny=3, dx=0.05, dy=0.05), dict(nx=90, ny=40, dx=-0.05, dy=0.05), dict(nx=45, ny=20, dx=0.5, dy=0.35),
                                  dict(nx=200, ny=5, dx=0.0, dy=0.05), dict(nx=1, ny=1, dx=0.05, dy=0.05), dict(nx=65, ny=9, dx=0.013, dy=-0.021)])
def test_lattice_odd_geometries(cb, ctx, geom):
    """Lattices with fewer rows than a unit (several headings inside one unit), decreasing x, coarse and
    sub-cell pitches, a degenerate pitch and a single pose."""
    rng = np.random.default_rng(35)
    data = mixed_map(rng, 320, 300, lethal=0.003)
    th = -math.pi + 2 * math.pi * np.arange(9) / 9
    x0 = 12.0 if geom["dx"] < 0 else 0.6
    desc = dict(x0=x0, dx=geom["dx"], nx=geom["nx"], y0=7.0 if geom["dy"] < 0 else 0.7, dy=geom["dy"], ny=geom["ny"], cs=np.stack([np.cos(th), np.sin(th)], 1))
    cmap = cb.Costmap(ctx, data, 0.05, 0.1, -0.2)
    gpu = cmap.area_is_free(RECT, cb.Lattice(**desc))
    cpu = po.polygon_batch(po.Costmap(data, 0.05, 0.1, -0.2), RECT, po.Lattice(**desc), threads=8)
    assert_same(gpu, cpu, what=str(geom))
    out = cb.Result(free_bits=np.zeros((len(cpu.is_free) + 31) // 32, dtype=np.uint32), not_ok=np.zeros(1, dtype=np.int64))
    cmap.area_is_free(RECT, cb.Lattice(**desc), out=out)
    np.testing.assert_array_equal(out.unpack_bits(len(cpu.is_free)), cpu.is_free)
    for name, op in (("accumulate_cost", po.OP_ACCUMULATE), ("max_cost", po.OP_MAX)):
        assert_same(getattr(cmap, f"area_{name}")(RECT, cb.Lattice(**desc)),
                    po.polygon_batch(po.Costmap(data, 0.05, 0.1, -0.2), RECT, po.Lattice(**desc), op=op, threads=8), True, f"{name} {geom}")


def test_lattice_many_distinct_shapes(cb, ctx):
    """Sub-cell pitches make the relative shape change from pose to pose: hundreds of distinct shapes per
    heading overflow the 64-slot shape cache, so poses spill to the todo list / general kernel; 5- and
    12-vertex polygons take the hashed-key path with verification."""
    rng = np.random.default_rng(36)
    data = mixed_map(rng, 200, 200, lethal=0.004)
    cmap = cb.Costmap(ctx, data, 0.05)
    omap = po.Costmap(data, 0.05)
    th = np.array([0.0, 0.7, -2.1])
    ang = 2 * math.pi * np.arange(12) / 12
    polys = [RECT, np.stack([0.45 * np.cos(ang), 0.45 * np.sin(ang)]), np.array([[0.4, 0.1, -0.4, -0.2, 0.3], [0.0, 0.35, 0.2, -0.3, -0.3]])]
    for poly in polys:
        desc = dict(x0=0.8, dx=0.0137, nx=300, y0=0.8, dy=0.0211, ny=40, cs=np.stack([np.cos(th), np.sin(th)], 1))
        gpu = cmap.area_is_free(poly, cb.Lattice(**desc))
        cpu = po.polygon_batch(omap, poly, po.Lattice(**desc), threads=8)
        assert_same(gpu, cpu, what=f"{poly.shape[1]} vertices")
        assert 0.02 < cpu.is_free.mean() < 0.98
        assert_same(cmap.area_accumulate_cost(poly, cb.Lattice(**desc)), po.polygon_batch(omap, poly, po.Lattice(**desc), op=po.OP_ACCUMULATE, threads=8), True)


def test_lattice_other_small_footprints(cb, ctx):
    """Triangles, the L-shape scaled down, a 1-cell polygon and a degenerate (open) outline on a lattice."""
    rng = np.random.default_rng(34)
    data = mixed_map(rng, 256, 256, lethal=0.004)
    cmap = cb.Costmap(ctx, data, 0.05)
    omap = po.Costmap(data, 0.05)
    th = -math.pi + 2 * math.pi * np.arange(12) / 12
    desc = dict(x0=0.2, dx=0.05, nx=230, y0=0.3, dy=0.05, ny=100, cs=np.stack([np.cos(th), np.sin(th)], 1))
    polys = [np.array([[0.4, -0.3, -0.3], [0.0, 0.25, -0.25]]), FOOTPRINTS[6] * 0.8, np.array([[0.01], [0.01]]),
             np.array([[0.0, 0.3], [0.0, 0.45]]), np.array([[0.6, 0.6, -0.6, -0.6], [0.05, -0.05, -0.05, 0.05]])]
    for poly in polys:
        gpu = cmap.area_is_free(poly, cb.Lattice(**desc))
        cpu = po.polygon_batch(omap, poly, po.Lattice(**desc), threads=8)
        assert_same(gpu, cpu, what=str(poly.tolist()))
        for name, op in (("accumulate_cost", po.OP_ACCUMULATE), ("max_cost", po.OP_MAX)):
            assert_same(getattr(cmap, f"area_{name}")(poly, cb.Lattice(**desc)), po.polygon_batch(omap, poly, po.Lattice(**desc), op=op, threads=8),
                        True, f"{name} {poly.tolist()}")


def test_lattice_non_convex_footprints(cb, ctx):
    """Non-convex footprints on a lattice: scan rows with two merged range pairs go through the four-range rasteriser
    of the shape cache (span table), rows with more through the general kernel - cusps, ties and all, against the
    oracle for the three folds, on a cell-pitch and on a sub-cell-pitch lattice."""
    from cover_b200 import workloads as wl
    rng = np.random.default_rng(91)
    data = mixed_map(rng, 320, 300, lethal=0.003)
    cmap = cb.Costmap(ctx, data, 0.05, -1.0, -0.5)
    omap = po.Costmap(data, 0.05, -1.0, -0.5)
    th = np.concatenate([-math.pi + 2 * math.pi * np.arange(10) / 10, [0.0, math.pi / 2]])
    cs = np.stack([np.cos(th), np.sin(th)], 1)
    notch = np.array([[-0.5, 0.5, 0.5, 0.15, 0.15, -0.15, -0.15, -0.5], [-0.4, -0.4, 0.4, 0.4, -0.1, -0.1, 0.4, 0.4]])  # a U
    star = np.stack([np.where(np.arange(16) % 2 == 0, 0.6, 0.22) * np.cos(2 * math.pi * np.arange(16) / 16),
                     np.where(np.arange(16) % 2 == 0, 0.6, 0.22) * np.sin(2 * math.pi * np.arange(16) / 16)])  # > 4 ranges on some rows
    polys = {"potato": wl.scaled(wl.potato(16), 0.6), "U": notch, "star": star, "L": FOOTPRINTS[6], "S x0.25": FOOTPRINTS[0] * 0.25,
             "TORU": FOOTPRINTS[5]}
    for pitch in (0.05, 0.0173):
        desc = dict(x0=-0.4, dx=pitch, nx=150, y0=0.2, dy=pitch, ny=60, cs=cs)
        for name, poly in polys.items():
            for op_name, op, has_cost in OPS:
                gpu = getattr(cmap, f"area_{op_name}")(poly, cb.Lattice(**desc))
                cpu = po.polygon_batch(omap, poly, po.Lattice(**desc), op=op, threads=8)
                assert_same(gpu, cpu, has_cost, f"{name} {op_name} pitch {pitch}")
            assert cpu.is_free.any() or name == "star"


def test_lattice_outlines(cb, ctx):
    """outline_generator checks over a lattice: verdict and maximum go through the shape cache with the outline's
    same-row runs as spans (any-reductions), accumulate_cost through the per-pose kernel; map border included."""
    from cover_b200 import workloads as wl
    rng = np.random.default_rng(93)
    data = mixed_map(rng, 300, 260, lethal=0.004)
    cmap = cb.Costmap(ctx, data, 0.05, 0.5, 0.25)
    omap = po.Costmap(data, 0.05, 0.5, 0.25)
    th = np.concatenate([-math.pi + 2 * math.pi * np.arange(9) / 9, [0.0, -math.pi / 2]])
    cs = np.stack([np.cos(th), np.sin(th)], 1)
    spike = np.array([[0.0, 0.5, 0.0, -0.02], [0.0, 0.0, 0.02, 0.6]])  # thin spikes: runs that turn around on a row
    polys = {"rect": RECT, "L": FOOTPRINTS[6], "TORU": FOOTPRINTS[5], "potato": wl.scaled(wl.potato(16), 0.6), "2 x rect": 2.0 * RECT,
             "spike": spike, "point": np.array([[0.01], [0.01]]), "segment": np.array([[0.0, 0.3], [0.0, 0.45]])}
    for pitch in (0.05, 0.0211):
        desc = dict(x0=0.2, dx=pitch, nx=170, y0=0.0, dy=pitch, ny=70, cs=cs)  # (the first rows / columns hang over the map border)
        for name, poly in polys.items():
            for op_name, op, has_cost in OPS:
                gpu = getattr(cmap, f"outline_{op_name}")(poly, cb.Lattice(**desc))
                cpu = po.polygon_batch(omap, poly, po.Lattice(**desc), shape=po.SHAPE_OUTLINE, op=op, threads=8)
                assert_same(gpu, cpu, has_cost, f"outline {name} {op_name} pitch {pitch}")
    # a sub-range that starts and ends inside a row
    desc = dict(x0=1.0, dx=0.05, nx=170, y0=1.0, dy=0.05, ny=70, cs=cs)
    gpu = cmap.outline_is_free(RECT, cb.Lattice(**desc), first=12_345, count=77_777)
    cpu = po.polygon_batch(omap, RECT, po.Lattice(**desc), shape=po.SHAPE_OUTLINE, first=12_345, count=77_777, threads=8)
    assert_same(gpu, cpu, False, "outline sub-range")


def test_lattice_big_footprints(cb, ctx):
    """Footprints up to ~70 scan rows on a lattice: pairs wider than 32 cells are split into several spans of the
    shape's span table (the 1.0 x 0.6 m rectangle on a 2 cm map, a 2.6 x 1.5 m rectangle, SOTO, a fat U at 5 cm);
    the W footprint is taller than the cache takes and goes through the general kernels as before."""
    rng = np.random.default_rng(92)
    th = np.concatenate([-math.pi + 2 * math.pi * np.arange(7) / 7, [0.0, math.pi / 2]])
    cs = np.stack([np.cos(th), np.sin(th)], 1)
    fat_u = 2.2 * np.array([[-0.5, 0.5, 0.5, 0.15, 0.15, -0.15, -0.15, -0.5], [-0.4, -0.4, 0.4, 0.4, -0.1, -0.1, 0.4, 0.4]])
    cases = [(0.02, RECT, "rect @ 2 cm"), (0.05, 2.6 * RECT, "2.6 x rect"), (0.05, FOOTPRINTS[4], "SOTO"), (0.05, fat_u, "fat U"),
             (0.05, FOOTPRINTS[1], "W")]
    for res, poly, name in cases:
        data = mixed_map(rng, 420, 380, lethal=0.0015)
        cmap = cb.Costmap(ctx, data, res, -1.0 * res / 0.05, -0.5 * res / 0.05)
        omap = po.Costmap(data, res, -1.0 * res / 0.05, -0.5 * res / 0.05)
        for pitch in (res, 0.37 * res):
            desc = dict(x0=-0.5 * res / 0.05, dx=pitch, nx=140, y0=0.3 * res / 0.05, dy=pitch, ny=40, cs=cs)
            for op_name, op, has_cost in OPS:
                gpu = getattr(cmap, f"area_{op_name}")(poly, cb.Lattice(**desc))
                cpu = po.polygon_batch(omap, poly, po.Lattice(**desc), op=op, threads=8)
                assert_same(gpu, cpu, has_cost, f"{name} {op_name} pitch {pitch}")


@pytest.mark.parametrize("idx", range(7))
@pytest.mark.parametrize("shape,prefix", [(po.SHAPE_AREA, "area"), (po.SHAPE_OUTLINE, "outline")])
def test_reference_footprints_all_ops(cb, ctx, idx, shape, prefix):
    """The reference's 7 test footprints (non-convex ones exercise cusps, ties and the list store)."""
    rng = np.random.default_rng(10 + idx)
    data = mixed_map(rng, 600, 500, lethal=0.0004)
    poses = poses_around(rng, 4000, -1.0, 27.0)
    cmap = cb.Costmap(ctx, data, 0.05, -0.5, 1.0)
    omap = po.Costmap(data, 0.05, -0.5, 1.0)
    for name, op, has_cost in OPS:
        gpu = getattr(cmap, f"{prefix}_{name}")(FOOTPRINTS[idx], poses)
        cpu = po.polygon_batch(omap, FOOTPRINTS[idx], poses, shape=shape, op=op, threads=8)
        assert_same(gpu, cpu, has_cost, f"{FOOTPRINT_NAMES[idx]} {prefix} {name}")


@pytest.mark.parametrize("res", [0.2, 0.1, 0.025])
def test_resolutions_and_small_scale(cb, ctx, res):
    rng = np.random.default_rng(20)
    data = mixed_map(rng, 300, 260, lethal=0.003)
    cmap = cb.Costmap(ctx, data, res)
    omap = po.Costmap(data, res)
    span = 300 * res
    for fp in (FOOTPRINTS[5], FOOTPRINTS[6], RECT, FOOTPRINTS[0] * 0.4):
        poses = poses_around(rng, 3000, 0.0, span)
        assert_same(cmap.area_accumulate_cost(fp, poses), po.polygon_batch(omap, fp, poses, op=po.OP_ACCUMULATE, threads=8), True)


def test_degenerate_polygons_and_statuses(cb, ctx):
    """Empty / 1-point / 2-point / collinear polygons; open outlines give LOGIC_ERROR (generators.cpp:246-249)."""
    rng = np.random.default_rng(30)
    data = mixed_map(rng, 200, 200, lethal=0.01)
    cmap = cb.Costmap(ctx, data, 0.05)
    omap = po.Costmap(data, 0.05)
    poses = poses_around(rng, 2000, 0.5, 9.5)
    polys = [
        np.zeros((2, 0)),
        np.array([[0.2], [0.1]]),
        np.array([[0.0, 0.0, 0.0], [0.0, 0.0, 0.0]]),
        np.array([[0.0, 0.4], [0.0, 0.0]]),
        np.array([[0.0, 0.3], [0.0, 0.45]]),
        np.array([[0.0, 0.3, 0.6], [0.0, 0.3, 0.6]]),
        np.array([[0.0, 0.05, 0.1, 0.05], [0.0, 0.05, 0.0, 0.05]]),
        np.array([[0.0, 0.3, 0.3, 0.0, 0.0, 0.3], [0.0, 0.0, 0.3, 0.3, 0.0, 0.0]]),
    ]
    saw_logic_error = False
    for poly in polys:
        for shape, prefix in ((po.SHAPE_AREA, "area"), (po.SHAPE_OUTLINE, "outline")):
            for name, op, has_cost in OPS:
                gpu = getattr(cmap, f"{prefix}_{name}")(poly, poses)
                cpu = po.polygon_batch(omap, poly, poses, shape=shape, op=op, threads=4)
                assert_same(gpu, cpu, has_cost, f"{poly.tolist()} {prefix} {name}")
                saw_logic_error |= bool((cpu.status == po.LOGIC_ERROR).any())
    assert saw_logic_error


def test_random_small_polygons_fuzz(cb, ctx):
    """Random 3..9-gons on a coarse grid: provokes every scan-line quirk through the full pipeline."""
    rng = np.random.default_rng(31)
    data = mixed_map(rng, 120, 120, lethal=0.01)
    cmap = cb.Costmap(ctx, data, 0.1)
    omap = po.Costmap(data, 0.1)
    poses = poses_around(rng, 300, 1.0, 11.0)
    for _ in range(120):
        n = int(rng.integers(3, 10))
        poly = rng.uniform(-0.9, 0.9, (2, n))
        gpu = cmap.area_accumulate_cost(poly, poses)
        cpu = po.polygon_batch(omap, poly, poses, op=po.OP_ACCUMULATE, threads=4)
        assert_same(gpu, cpu, True, str(poly.tolist()))


def test_no_pose_overload_and_coord_range(cb, ctx):
    """is_free(polygon, map) (costmap.cpp:71-76) and the COORD_RANGE status."""
    data = np.zeros((100, 100), dtype=np.uint8)
    data[40, 40] = 254
    cmap = cb.Costmap(ctx, data, 0.1, -2.0, -3.0)
    omap = po.Costmap(data, 0.1, -2.0, -3.0)
    for poly in (FOOTPRINTS[6] * 0.5 + 1.0, FOOTPRINTS[6] * 0.5 + np.array([[2.0], [1.0]]), RECT + 50.0):
        gpu, cpu = cmap.area_is_free(poly, None), po.polygon_batch(omap, poly, None)
        assert_same(gpu, cpu)
        gpu, cpu = cmap.outline_accumulate_cost(poly, None), po.polygon_batch(omap, poly, None, shape=po.SHAPE_OUTLINE, op=po.OP_ACCUMULATE)
        assert_same(gpu, cpu, True)
    far = np.array([[1.0, 0.0, 1e7, 0.0], [1.0, 0.0, 5.0, 5.0], [float("nan"), 0.0, 1.0, 1.0]])
    gpu, cpu = cmap.area_accumulate_cost(RECT, far), po.polygon_batch(omap, RECT, far, op=po.OP_ACCUMULATE)
    assert cpu.status.tolist() == [po.COORD_RANGE, po.OK, po.COORD_RANGE]
    assert_same(gpu, cpu, True)


def test_golden_vectors_through_the_gpu(cb, ctx):
    """test/costmap.cpp:106-194 known answers, evaluated by the CUDA kernels."""
    data = np.zeros((20, 10), dtype=np.uint8)
    data[5, 5], data[6, 5], data[7, 5] = 254, 255, 253
    cmap = cb.Costmap(ctx, data, 1.0)
    assert cmap.ray_is_free([(4, 0, 4, 9)]).is_free[0] == 1
    assert cmap.ray_is_free([(5, 0, 5, 9)]).is_free[0] == 0
    assert cmap.ray_is_free([(-4, -4, -4, 4)]).is_free[0] == 0
    assert cmap.ray_is_free([(-4, -4, -4, 4)], offset=(4, 4)).is_free[0] == 1
    tens = cb.Costmap(ctx, np.full((20, 10), 10, dtype=np.uint8), 1.0)
    sparse = np.array([(0, 0), (9, 0), (9, 19), (0, 19)], dtype=np.int32)
    area = cb.DiscretePolygon(ctx, po.area_cells([po.outline_cells(sparse)]))
    assert area.accumulate_cost(tens, [(0, 0)]).cost[0] == 2000
    shifted = cb.DiscretePolygon(ctx, po.area_cells([po.outline_cells(sparse - np.array([4, 4], dtype=np.int32))]))
    assert shifted.accumulate_cost(tens, [(0, 0)]).cost[0] == -3
    assert shifted.accumulate_cost(tens, [(4, 4)]).cost[0] == 2000
    assert tens.ray_accumulate_cost([(-1, -1, 5, 5)]).cost[0] == -3
    # box polygon covering the whole map through the polygon entry: sum = 10 * 200
    box = np.array([[0.5, 9.5, 9.5, 0.5], [0.5, 0.5, 19.5, 19.5]])
    assert tens.area_accumulate_cost(box, None).cost[0] == 2000


def test_rays_cells_and_footprint_bank(cb, ctx):
    rng = np.random.default_rng(40)
    data = mixed_map(rng, 256, 192, lethal=0.01)
    cmap = cb.Costmap(ctx, data, 0.05)
    omap = po.Costmap(data, 0.05)
    segs = rng.integers(-10, 270, (50_000, 4)).astype(np.int32)
    for name, op, has_cost in OPS:
        for off in (None, (7, -3)):
            gpu = getattr(cmap, f"ray_{name}")(segs, off)
            cpu = po.rays_batch(omap, segs, off, op, threads=8)
            assert_same(gpu, cpu, has_cost, f"ray {name} {off}")
    cells = po.area_cells([po.outline_cells(po.to_discrete(0.05, FOOTPRINTS[5]))])
    offs = rng.integers(-20, 280, (40_000, 2)).astype(np.int32)
    dp = cb.DiscretePolygon(ctx, cells)
    for name, op, has_cost in OPS:
        assert_same(getattr(dp, name)(cmap, offs), po.cells_batch(omap, cells, offs, op, threads=8), has_cost, f"cells {name}")

    toru = FOOTPRINTS[5]
    outlines, areas = [toru * 0.45], [toru, RECT * 0.4 + np.array([[0.5], [0.0]])]
    bank = []
    for k in range(12):
        th = 2 * math.pi * k / 12
        R = np.array([[math.cos(th), -math.sin(th)], [math.sin(th), math.cos(th)]])
        bank.append(po.make_discrete_footprint(0.05, [R @ o for o in outlines], [R @ a for a in areas]))
    which = rng.integers(0, 12, len(offs)).astype(np.int32)
    fb = cb.FootprintBank(ctx, bank)
    gpu = fb.is_free(cmap, offs, which)
    cpu = po.discrete_footprint_batch(omap, bank, offs, which, threads=8)
    np.testing.assert_array_equal(gpu.is_free, cpu.is_free)
    assert 0 < cpu.is_free.mean() < 1
    np.testing.assert_array_equal(fb.is_free(cmap, offs).is_free, po.discrete_footprint_batch(omap, bank, offs, None, threads=8).is_free)


def test_split_footprint_pose_form(cb, ctx):
    """continuous_footprint::is_free(pose, map) with hand-made splits, incl. test/footprints.cpp:244-305."""
    rng = np.random.default_rng(50)
    data = np.zeros((200, 100), dtype=np.uint8)
    box = np.array([[0, 1, 1, 0], [0, 0, 2, 2]], dtype=np.float64) - np.array([[0.5], [1.0]])
    ident = np.array([[1.0, 0.0, 0.0, 0.0]])

    def both(d, outlines, areas, poses):
        g = cb.SplitFootprint(ctx, outlines, areas).is_free(cb.Costmap(ctx, d, 0.1, -5.0, -10.0), poses)
        c = po.continuous_footprint_batch(po.Costmap(d, 0.1, -5.0, -10.0), outlines, areas, poses)
        np.testing.assert_array_equal(g.is_free, c.is_free)
        np.testing.assert_array_equal(g.status, c.status)
        return g.is_free

    assert both(data, [box], [], ident)[0] == 1
    data[100, 50] = 254
    assert both(data, [box], [], ident)[0] == 1
    data[90, 45] = 253
    assert both(data, [box], [], ident)[0] == 0
    assert both(data, [], [box], ident)[0] == 0
    data[100, 50] = 253
    assert both(data, [], [box], ident)[0] == 1
    assert both(data, [box * 0.5], [box], np.array([[1.0, 0.0, 5.0, 10.0]]))[0] == 0

    big = mixed_map(rng, 500, 400, lethal=0.0008)
    toru = FOOTPRINTS[5]
    outlines = [toru * 0.4, toru * 0.2 + np.array([[0.3], [0.0]])]
    areas = [toru, FOOTPRINTS[6] * 0.8, RECT * 0.5 + np.array([[-0.2], [0.1]])]
    poses = poses_around(rng, 30_000, -1.0, 24.0)
    g = cb.SplitFootprint(ctx, outlines, areas).is_free(cb.Costmap(ctx, big, 0.05, -1.0, 0.5), poses)
    c = po.continuous_footprint_batch(po.Costmap(big, 0.05, -1.0, 0.5), outlines, areas, poses, threads=8)
    np.testing.assert_array_equal(g.status, c.status)
    np.testing.assert_array_equal(g.is_free, c.is_free)
    assert 0.02 < c.is_free.mean() < 0.98


@pytest.mark.parametrize("env", [{}, {"COVER_B200_TODO_CAP": "64"}, {"COVER_B200_NO_SPLIT_LOCKSTEP": "1"}, {"COVER_B200_PLANES_MIN_BATCH": "0"}])
def test_split_footprint_lockstep_and_deferrals(cb, monkeypatch, env):
    """k_split_lockstep + what it leaves to k_split: areas with more than four ranges on a row (M, Spiral), an open
    polyline as an outline, poses over the map border and out of coordinate range, the not-OK counter, bit-packed
    verdicts; the same with an overflowing todo list, without the lock-step pass, and with the bit planes attached."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    c = cb.Context(0)
    rng = np.random.default_rng(52)
    big = mixed_map(rng, 420, 400, lethal=0.001)
    cmap, omap = cb.Costmap(c, big, 0.05, -1.0, 0.5), po.Costmap(big, 0.05, -1.0, 0.5)
    poses = poses_around(rng, 40_000, -2.0, 22.0)
    poses[::97, 2] = 3e9  # coordinates out of range
    cases = [
        ([RECT * 0.4], [FOOTPRINTS[2] * 0.3, FOOTPRINTS[3] * 0.25]),                     # M and Spiral as areas
        ([np.array([[0.0, 0.6, 0.9], [0.0, 0.1, -0.4]]), RECT * 0.3], [FOOTPRINTS[6]]),  # an open polyline outline, L area
        ([], [FOOTPRINTS[0] * 0.35, RECT]),
        ([FOOTPRINTS[5] * 0.5], []),
    ]
    for outlines, areas in cases:
        sf = cb.SplitFootprint(c, outlines, areas)
        g = sf.is_free(cmap, poses)
        r = po.continuous_footprint_batch(omap, outlines, areas, poses, threads=8)
        np.testing.assert_array_equal(g.status, r.status, err_msg=str(env))
        np.testing.assert_array_equal(g.is_free, r.is_free, err_msg=str(env))
        out = cb.Result(free_bits=np.zeros((len(poses) + 31) // 32, dtype=np.uint32), not_ok=np.zeros(1, dtype=np.int64))
        sf.is_free(cmap, poses, out=out)
        np.testing.assert_array_equal(out.unpack_bits(len(poses)), r.is_free & (r.status == 0))
        assert int(out.not_ok[0]) == int((r.status != 0).sum()) > 0
    c.close()


@pytest.fixture()
def ctx_planes(cb, monkeypatch):
    """A context that builds the lethal / inscribed bit planes for EVERY batch (the default only does for >= 32768)."""
    monkeypatch.setenv("COVER_B200_PLANES_MIN_BATCH", "0")
    c = cb.Context(0)
    yield c
    c.close()


def test_span_scans_over_bit_planes(cb, ctx_planes):
    """Verdict kernels answer a row span from two words of the bit planes (span_any_bits) instead of scanning its cost
    bytes: every span-scanning path against the oracle with the planes forced on - random-pose areas of all
    reference footprints (narrow and > 32-cell spans, map border), cached cells, the footprint bank (253 | 254
    planes for outline spans), the split footprint - and across map updates (the planes must follow the image)."""
    ctx = ctx_planes
    rng = np.random.default_rng(77)
    data = mixed_map(rng, 700, 500, lethal=0.0015)
    cmap = cb.Costmap(ctx, data, 0.05)
    omap = po.Costmap(data, 0.05)
    poses = poses_around(rng, 6000, -1.0, 36.0)  # (poses hang over the map border on all sides)
    for idx in range(7):
        for scale in (1.0, 2.7):  # 2.7: spans wider than 32 cells
            poly = FOOTPRINTS[idx] * scale
            gpu = cmap.area_is_free(poly, poses)
            cpu = po.polygon_batch(omap, poly, poses, po.SHAPE_AREA, po.OP_IS_FREE, threads=8)
            assert_same(gpu, cpu, False, f"{FOOTPRINT_NAMES[idx]} x{scale}")
    assert ctx.launch_count > 0
    # the image changes: full upload, window upload, inflation - the verdicts must follow
    data2 = mixed_map(rng, 700, 500, lethal=0.004)
    cmap.upload(data2)
    omap2 = po.Costmap(data2, 0.05)
    assert_same(cmap.area_is_free(RECT, poses), po.polygon_batch(omap2, RECT, poses, po.SHAPE_AREA, po.OP_IS_FREE, threads=8), False, "after upload")
    data3 = data2.copy()
    data3[100:300, 200:450] = mixed_map(rng, 250, 200, lethal=0.02)
    cmap.upload_window(data3, 200, 100, 250, 200)
    omap3 = po.Costmap(data3, 0.05)
    assert_same(cmap.area_is_free(RECT, poses), po.polygon_batch(omap3, RECT, poses, po.SHAPE_AREA, po.OP_IS_FREE, threads=8), False, "after window upload")
    # cached cells and the footprint bank
    cells = po.area_cells([po.outline_cells(po.to_discrete(0.05, FOOTPRINTS[5] * 1.8))])
    offs = rng.integers(-30, 720, (30_000, 2)).astype(np.int32)
    assert_same(cb.DiscretePolygon(ctx, cells).is_free(cmap, offs), po.cells_batch(omap3, cells, offs, po.OP_IS_FREE, threads=8), False, "cells")
    toru = FOOTPRINTS[5]
    outlines, areas = [toru * 0.45], [toru * 1.6, RECT * 0.4 + np.array([[0.5], [0.0]])]
    bank = []
    for k in range(8):
        th = 2 * math.pi * k / 8
        R = np.array([[math.cos(th), -math.sin(th)], [math.sin(th), math.cos(th)]])
        bank.append(po.make_discrete_footprint(0.05, [R @ o for o in outlines], [R @ a for a in areas]))
    which = rng.integers(0, 8, len(offs)).astype(np.int32)
    got = cb.FootprintBank(ctx, bank).is_free(cmap, offs, which)
    want = po.discrete_footprint_batch(omap3, bank, offs, which, threads=8)
    np.testing.assert_array_equal(got.is_free, want.is_free)
    assert want.is_free.any() and not want.is_free.all()
    # split footprint (areas over the planes, outlines cell by cell)
    sp_out = [toru * 0.4, toru * 0.2 + np.array([[0.3], [0.0]])]
    sp_area = [toru * 1.5, FOOTPRINTS[6] * 0.8, RECT * 0.5 + np.array([[-0.2], [0.1]])]
    g = cb.SplitFootprint(ctx, sp_out, sp_area).is_free(cmap, poses)
    c = po.continuous_footprint_batch(omap3, sp_out, sp_area, poses, threads=8)
    np.testing.assert_array_equal(g.status, c.status)
    np.testing.assert_array_equal(g.is_free, c.is_free)


def test_device_resident_inputs_and_outputs(cb, ctx):
    """Same numbers when poses/results stay on the device (the bench's `value` path)."""
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(60)
    data = mixed_map(rng, 512, 512)
    poses = poses_around(rng, 20_000, 0.0, 25.6)
    cmap = cb.Costmap(ctx, torch.from_numpy(data).cuda(), 0.05)
    host = cmap.area_accumulate_cost(RECT, poses)
    dposes = torch.from_numpy(poses).cuda()
    out = cb.Result(torch.empty(len(poses), dtype=torch.uint8, device="cuda"), torch.empty(len(poses), dtype=torch.float64, device="cuda"),
                    torch.empty(len(poses), dtype=torch.uint8, device="cuda"))
    cmap.area_accumulate_cost(RECT, dposes, out=out)
    ctx.synchronize()
    np.testing.assert_array_equal(out.is_free.cpu().numpy(), host.is_free)
    np.testing.assert_array_equal(out.cost.cpu().numpy(), host.cost)
    cpu = po.polygon_batch(po.Costmap(data, 0.05), RECT, poses, op=po.OP_ACCUMULATE, threads=8)
    assert_same(host, cpu, True)


def test_map_window_update(cb, ctx):
    rng = np.random.default_rng(70)
    data = mixed_map(rng, 300, 300, lethal=0.0)
    cmap = cb.Costmap(ctx, data, 0.05)
    poses = poses_around(rng, 5000, 1.0, 14.0)
    assert cmap.area_is_free(RECT, poses).is_free.all()
    data[100:140, 120:170] = 254
    cmap.upload_window(data, 120, 100, 50, 40)
    assert_same(cmap.area_is_free(RECT, poses), po.polygon_batch(po.Costmap(data, 0.05), RECT, poses, threads=8))
    # with the bit planes and the prefix rows built (lattice sweeps): a window update repairs their rows in place
    th = np.array([0.3, -1.2, 2.0])
    desc = dict(x0=1.0, dx=0.05, nx=200, y0=1.0, dy=0.05, ny=200, cs=np.stack([np.cos(th), np.sin(th)], 1))
    for rep in range(3):
        for name, op, has_cost in OPS[:2]:
            assert_same(getattr(cmap, f"area_{name}")(RECT, cb.Lattice(**desc)), po.polygon_batch(po.Costmap(data, 0.05), RECT, po.Lattice(**desc), op=op, threads=8),
                        has_cost, f"lattice {name} after {rep} window updates")
        x0, y0 = int(rng.integers(0, 200)), int(rng.integers(0, 220))
        data[y0:y0 + 70, x0:x0 + 90] = mixed_map(rng, 90, 70, lethal=0.01)
        cmap.upload_window(data, x0, y0, 90, 70)


def test_full_size_lattice_properties(cb, ctx):
    """BASELINE config 4 at FULL size (4000 x 4000 map, 1180 x 1180 x 72 lattice = 1.0e8 poses): sampled
    oracle comparison + size-independent properties (idempotence, verdict <-> max_cost consistency)."""
    from cover_b200 import workloads as wl
    data = wl.inflated_costmap(4000, 4000, 0.05, seed=3)
    cmap = cb.Costmap(ctx, data, 0.05)
    desc = wl.dwa_lattice()
    lat = cb.Lattice(**desc)
    first = cmap.area_is_free(RECT, lat)
    assert len(first.is_free) == 1180 * 1180 * 72
    assert (first.status == 0).all()
    again = cmap.area_is_free(RECT, lat)
    np.testing.assert_array_equal(first.is_free, again.is_free)
    frac = first.is_free.mean()
    assert 0.5 < frac < 0.999
    # sampled parity against the oracle
    rng = np.random.default_rng(4)
    olat = po.Lattice(**desc)
    omap = po.Costmap(data, 0.05)
    for start in rng.integers(0, lat.count - 40_000, 6).tolist():
        cpu = po.polygon_batch(omap, RECT, olat, first=start, count=40_000, threads=8)
        np.testing.assert_array_equal(first.is_free[start:start + 40_000], cpu.is_free)
    # verdict <-> max consistency on a slab: no 255 in this map, so is_free == (0 <= max < 254)
    n = 3_000_000
    mx = cmap.area_max_cost(RECT, lat, first=5_000_000, count=n)
    np.testing.assert_array_equal(first.is_free[5_000_000:5_000_000 + n], ((mx.cost >= 0) & (mx.cost < 254)).astype(np.uint8))


def test_async_banded_upload_is_ordered_before_every_reader(cb, ctx):
    """cover_map_upload_async: the image travels as row bands on a second stream (the bands under the last sweep's
    window first) and a lattice sweep starts once the bands under ITS window have landed.  Alternating two
    different costmaps and several windows, every check must see exactly the image uploaded just before it -
    compared with the synchronous upload on a second map object."""
    import torch
    from cover_b200 import workloads as wl
    images = [wl.inflated_costmap(4000, 4000, 0.05, seed=3), wl.inflated_costmap(4000, 4000, 0.05, seed=11, obstacle_p=6e-4)]
    pinned = [torch.from_numpy(im).pin_memory().numpy() for im in images]
    sync_map = cb.Costmap(ctx, images[0], 0.05)
    amap = cb.Costmap(ctx, images[0], 0.05)
    lattices = [cb.Lattice(**wl.dwa_lattice(n_theta=24)),                                        # low window
                cb.Lattice(**wl.dwa_lattice(nx=600, ny=500, n_theta=12, x0=100.0, y0=160.0)),  # high window: the hint changes
                cb.Lattice(**dict(wl.dwa_lattice(nx=800, ny=700, n_theta=10), y0=120.0, dy=-0.05)),  # negative row pitch
                cb.Lattice(**wl.dwa_lattice(nx=900, ny=2000, n_theta=8, x0=5.0, y0=-2.0, pitch=0.11)),  # leaves the map at both ends
                cb.Lattice(**dict(wl.dwa_lattice(nx=300, ny=300, n_theta=4), cs=2.5 * wl.heading_table(4)))]  # |(c,s)| != 1: longer reach
    expected = {}
    for li, lat in enumerate(lattices):
        for mi in (0, 1):
            sync_map.upload(images[mi])
            expected[li, mi] = sync_map.area_is_free(RECT, lat).is_free
        assert not np.array_equal(expected[li, 0], expected[li, 1])
    for rep in range(2):
        for li, lat in enumerate(lattices):
            for mi in (1, 0):
                amap.upload_async(pinned[mi])
                got = amap.area_is_free(RECT, lat)
                assert (got.status == 0).all()
                np.testing.assert_array_equal(got.is_free, expected[li, mi], err_msg=f"lattice {li} image {mi} rep {rep}")
    # two sweeps over different windows after ONE upload: the second needs bands the first did not wait for
    amap.upload_async(pinned[1])
    np.testing.assert_array_equal(amap.area_is_free(RECT, lattices[0]).is_free, expected[0, 1])
    np.testing.assert_array_equal(amap.area_is_free(RECT, lattices[1]).is_free, expected[1, 1])
    # a sub-range of the lattice, a value fold, a footprint too big for the lattice kernel, an outline
    lat = lattices[0]
    amap.upload_async(pinned[0])
    part = amap.area_is_free(RECT, lat, first=2_345_678, count=20_000_001)
    np.testing.assert_array_equal(part.is_free, expected[0, 0][2_345_678:2_345_678 + 20_000_001])
    big = 3.0 * RECT
    small_lat = cb.Lattice(**wl.dwa_lattice(nx=400, ny=300, n_theta=5, y0=150.0))
    for name, poly in (("area_accumulate_cost", RECT), ("area_is_free", big), ("outline_max_cost", big)):
        sync_map.upload(images[1])
        want = getattr(sync_map, name)(poly, small_lat)
        amap.upload_async(pinned[0])
        amap.area_is_free(RECT, lattices[0])  # moves the hint away from small_lat's window
        amap.upload_async(pinned[1])
        got = getattr(amap, name)(poly, small_lat)
        np.testing.assert_array_equal(got.is_free, want.is_free, err_msg=name)
        if want.cost is not None:
            np.testing.assert_array_equal(got.cost, want.cost, err_msg=name)
    # every other reader of the map waits for the whole image
    rng = np.random.default_rng(5)
    poses = poses_around(rng, 200_000, 2.0, 198.0)
    for mi in (1, 0):
        sync_map.upload(images[mi])
        amap.upload_async(pinned[mi])
        np.testing.assert_array_equal(amap.area_is_free(RECT, poses).is_free, sync_map.area_is_free(RECT, poses).is_free)
        amap.upload_async(pinned[mi])
        np.testing.assert_array_equal(amap.outline_accumulate_cost(RECT, poses).cost, sync_map.outline_accumulate_cost(RECT, poses).cost)
        amap.upload_async(pinned[mi])
        np.testing.assert_array_equal(amap.download(), images[mi])
    # the image buffer may be overwritten as soon as a check has returned its results to the host - also when that
    # check only waited for the bands under its window (the rest of the copy must have landed by then)
    scratch = torch.from_numpy(images[1].copy()).pin_memory().numpy()
    for _ in range(5):
        scratch[:] = images[1]
        amap.upload_async(scratch)
        small = amap.area_is_free(RECT, cb.Lattice(**wl.dwa_lattice(nx=200, ny=64, n_theta=4, x0=30.0, y0=6.0)))  # touches the first bands only
        scratch[:] = 254  # the caller refills its buffer for the next cycle
        assert small.is_free.any()
        np.testing.assert_array_equal(amap.download(), images[1])
    # a synchronous upload right after an asynchronous one wins
    amap.upload_async(pinned[1])
    amap.upload(images[0])
    np.testing.assert_array_equal(amap.download(), images[0])


def test_bit_packed_verdicts_and_not_ok_counter(cb, ctx):
    """free_bits / not_ok (cover_results) agree with the byte arrays on every kernel family."""
    rng = np.random.default_rng(80)
    data = mixed_map(rng, 400, 400, lethal=0.002)
    cmap = cb.Costmap(ctx, data, 0.05)

    def packed(n):
        return cb.Result(free_bits=np.full((n + 31) // 32, 0xFFFFFFFF, dtype=np.uint32), not_ok=np.full(1, -7, dtype=np.int64))

    # lattice kernel (unaligned sub-range), thread-per-pose kernel, list kernel, outline, rays
    th = -math.pi + 2 * math.pi * np.arange(5) / 5
    lat = cb.Lattice(0.4, 0.05, 333, 0.4, 0.05, 61, np.stack([np.cos(th), np.sin(th)], 1))
    ref = cmap.area_is_free(RECT, lat, first=777, count=90_001)
    out = packed(90_001)
    cmap.area_is_free(RECT, lat, first=777, count=90_001, out=out)
    np.testing.assert_array_equal(out.unpack_bits(90_001), ref.is_free)
    assert out.not_ok[0] == 0
    poses = poses_around(rng, 33_333, -0.5, 20.5)
    for poly in (RECT, FOOTPRINTS[0] * 0.3, np.array([[0.0, 0.3], [0.0, 0.45]])):
        for fn in (cmap.area_is_free, cmap.area_accumulate_cost, cmap.outline_is_free):
            ref = fn(poly, poses)
            out = packed(len(poses))
            fn(poly, poses, out=out)
            np.testing.assert_array_equal(out.unpack_bits(len(poses)), ref.is_free)
            assert out.not_ok[0] == int((ref.status != 0).sum())
    bad = np.array([[1.0, 0.0, 1e7, 0.0], [1.0, 0.0, 5.0, 5.0], [float("nan"), 0.0, 1.0, 1.0]])
    out = packed(3)
    cmap.area_is_free(RECT, bad, out=out)
    assert out.not_ok[0] == 2
    segs = rng.integers(-10, 410, (10_001, 4)).astype(np.int32)
    ref = cmap.ray_is_free(segs)
    out = packed(len(segs))
    cmap.ray_is_free(segs, out=out)
    np.testing.assert_array_equal(out.unpack_bits(len(segs)), ref.is_free)


def test_cpp_host_layer_runs_the_reference_unit_tests(cb, tmp_path):
    """tests/cpp/host_api_test.cpp: test/costmap.cpp + test/footprints.cpp known answers through
    include/cover_b200.hpp -> C ABI -> CUDA kernels."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.dirname(cb.library_path())
    exe = tmp_path / "host_api_test"
    build = subprocess.run(["g++", "-std=c++17", "-O1", "-I", os.path.join(root, "include"), os.path.join(root, "tests", "cpp", "host_api_test.cpp"),
                            "-o", str(exe), "-L", libdir, "-lcover_b200", f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
    assert build.returncode == 0, build.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert run.returncode == 0, run.stdout + run.stderr


def test_cpp_grid_split_matches_python_and_is_safe(cb, ctx, tmp_path):
    """tests/cpp/split_bank_test.cpp: the C++ host layer's to_area / split_cells / heading_bank (the Boost-free
    `continuous_footprint(radius, polygon)` of SURVEY 8f-1) must produce the same outline / area cells as
    cover_b200/split.py for every heading, and passes the reference's safety regression in C++."""
    import os
    import subprocess
    from cover_b200 import split, workloads as wl
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.dirname(cb.library_path())
    exe = tmp_path / "split_bank_test"
    build = subprocess.run(["g++", "-std=c++17", "-O1", "-I", os.path.join(root, "include"), os.path.join(root, "tests", "cpp", "split_bank_test.cpp"),
                            "-o", str(exe), "-L", libdir, "-lcover_b200", f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
    assert build.returncode == 0, build.stderr
    res, radius, n_headings = 0.05, 0.2, 12
    polygon = wl.potato(48)
    cs = wl.heading_table(n_headings)
    with open(tmp_path / "in.txt", "w") as f:
        f.write(f"{res!r} {radius!r}\n{polygon.shape[1]}\n")
        for x, y in polygon.T:
            f.write(f"{float(x)!r} {float(y)!r}\n")
        f.write(f"{n_headings}\n")
        for c, s_ in cs:
            f.write(f"{float(c)!r} {float(s_)!r}\n")
    run = subprocess.run([str(exe), str(tmp_path / "in.txt"), str(tmp_path / "out.txt")], capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, run.stdout + run.stderr
    # the same bank from the Python host layer (same map geometry as the C++ program: 240 x 240, origin at the centre)
    cmap = cb.Costmap(ctx, np.zeros((240, 240), dtype=np.uint8), res, -120 * res, -120 * res)
    poses = np.concatenate([cs, np.full((n_headings, 1), cmap.origin_x), np.full((n_headings, 1), cmap.origin_y)], axis=1)
    cells, starts, status = cmap.area_cells(polygon, poses)
    assert (status == 0).all()
    lines = open(tmp_path / "out.txt").read().split("\n")
    pos = 0
    for k in range(n_headings):
        tag, kk, n_out, n_area = lines[pos].split()
        assert tag == "H" and int(kk) == k
        n_out, n_area = int(n_out), int(n_area)
        got = np.array([ln.split() for ln in lines[pos + 1:pos + 1 + n_out + n_area]], dtype=np.int32).reshape(-1, 2)
        pos += 1 + n_out + n_area
        want_out, want_area = split.split_cells(cells[starts[k]:starts[k + 1]], res, radius)
        assert n_out == len(want_out) > 0 and n_area == len(want_area)

        def canon(a):
            return a[np.lexsort((a[:, 0], a[:, 1]))]
        np.testing.assert_array_equal(canon(got[:n_out]), canon(want_out), err_msg=f"outline cells, heading {k}")
        np.testing.assert_array_equal(canon(got[n_out:]), canon(want_area), err_msg=f"area cells, heading {k}")


def test_list_kernel_still_exact(cb, monkeypatch):
    """k_area_list (per-thread sorted lists in global scratch) is only used for polygons with more than 256
    vertices now; force it for ordinary polygons so that it stays covered."""
    monkeypatch.setenv("COVER_B200_FORCE_LIST_KERNEL", "1")
    c = cb.Context(0)
    try:
        rng = np.random.default_rng(90)
        data = mixed_map(rng, 500, 400, lethal=0.0005)
        cmap = cb.Costmap(c, data, 0.05)
        omap = po.Costmap(data, 0.05)
        poses = poses_around(rng, 2500, -1.0, 24.0)
        for idx in (0, 3, 6):
            for name, op, has_cost in OPS:
                gpu = getattr(cmap, f"area_{name}")(FOOTPRINTS[idx], poses)
                cpu = po.polygon_batch(omap, FOOTPRINTS[idx], poses, op=op, threads=8)
                assert_same(gpu, cpu, has_cost, f"list kernel {FOOTPRINT_NAMES[idx]} {name}")
    finally:
        c.close()


def test_many_vertex_polygons(cb, ctx):
    """64-, 256- and 300-vertex polygons (the last one exceeds the warp-per-pose kernel's vertex table)."""
    rng = np.random.default_rng(91)
    data = mixed_map(rng, 400, 400, lethal=0.001)
    cmap = cb.Costmap(ctx, data, 0.05)
    omap = po.Costmap(data, 0.05)
    poses = poses_around(rng, 1500, 0.0, 20.0)
    from cover_b200 import workloads as wl
    for nv in (64, 256, 300):
        poly = wl.potato(nv) * 1.7
        assert_same(cmap.area_accumulate_cost(poly, poses), po.polygon_batch(omap, poly, poses, op=po.OP_ACCUMULATE, threads=8), True, f"potato {nv}")
        star = poly * (1.0 + 0.35 * np.cos(9 * 2 * math.pi * np.arange(nv) / nv))
        assert_same(cmap.area_is_free(star, poses), po.polygon_batch(omap, star, poses, threads=8), False, f"star {nv}")


def _sparse_cells(poly, pose, res, ox, oy):
    """Path-A discretisation in numpy (every operation individually rounded, like the reference)."""
    c, s, tx, ty = pose
    px, py = np.asarray(poly, dtype=np.float64)
    qx = (tx + ((c * px) + ((-s) * py))) - ox
    qy = (ty + ((s * px) + (c * py))) - oy
    inv = 1.0 / res
    return np.stack([np.trunc(qx * inv), np.trunc(qy * inv)], 1).astype(np.int32)


def test_expand_matches_to_area_and_to_outline(cb, ctx):
    """cover_area_expand / cover_outline_expand == to_area / to_outline (expand.hpp:68-80) cell for cell, in
    generator order, for every reference footprint and for degenerate polygons."""
    rng = np.random.default_rng(95)
    cmap = cb.Costmap(ctx, np.zeros((50, 50), dtype=np.uint8), 0.05, -0.4, 0.3)
    poses = poses_around(rng, 60, -3.0, 3.0)
    polys = list(FOOTPRINTS) + [RECT, np.array([[0.0, 0.3], [0.0, 0.45]]), np.array([[0.1], [0.1]]), np.array([[0.0, 0.4], [0.0, 0.0]])]
    for poly in polys:
        a_cells, a_starts, a_status = cmap.area_cells(poly, poses)
        o_cells, o_starts, o_status = cmap.outline_cells(poly, poses)
        assert (o_status == 0).all()
        for i, pose in enumerate(poses):
            dense = po.outline_cells(_sparse_cells(poly, pose, 0.05, -0.4, 0.3))
            np.testing.assert_array_equal(o_cells[o_starts[i]:o_starts[i + 1]], dense)
            try:
                expect = po.area_cells([dense])
                assert a_status[i] == 0
                np.testing.assert_array_equal(a_cells[a_starts[i]:a_starts[i + 1]], expect)
            except RuntimeError:
                assert a_status[i] == po.LOGIC_ERROR and a_starts[i + 1] == a_starts[i]


@pytest.mark.parametrize("idx", [4, 5])
def test_grid_split_safety_regression(cb, ctx, idx):
    """The reference's safety property (test/footprints.cpp:205-242,351-391) for the grid split: with an
    inflation layer whose inscribed radius equals the split radius, a lethal obstacle on ANY cell of the
    footprint makes the split check fail (the reference tolerates 2 misses, this construction allows 0),
    and an empty map is free."""
    from cover_b200 import split, workloads as wl
    res, radius = 0.05, 0.2
    free_map = np.zeros((200, 200), dtype=np.uint8)
    cmap = cb.Costmap(ctx, free_map, res, -5.0, -5.0)
    bank, cs = split.heading_bank(cmap, FOOTPRINTS[idx], radius, 8)
    assert all(len(o) > 0 for o, _ in bank)
    fb = cb.FootprintBank(ctx, bank)
    centre = np.array([[100, 100]], dtype=np.int32)
    rad = int(math.ceil(radius / res))
    yy, xx = np.mgrid[-rad:rad + 1, -rad:rad + 1]
    disk = (np.hypot(xx, yy) * res <= radius)
    for k in (0, 3, 5):
        which = np.array([k], dtype=np.int32)
        assert fb.is_free(cmap, centre, which).is_free[0] == 1
        cells, starts, _ = cmap.area_cells(FOOTPRINTS[idx], np.array([[cs[k, 0], cs[k, 1], -5.0, -5.0]]))
        misses = 0
        for (cx, cy) in cells[::7]:  # every 7th footprint cell keeps the test short
            data = np.zeros((200, 200), dtype=np.uint8)
            x, y = cx + 100, cy + 100
            view = data[y - rad:y + rad + 1, x - rad:x + rad + 1]
            view[disk] = 253
            data[y, x] = 254
            cmap.upload(data)
            misses += int(fb.is_free(cmap, centre, which).is_free[0])
        cmap.upload(free_map)
        assert misses == 0
    # fewer cells are scanned than by the full-area check
    full = sum(len(cmap.area_cells(FOOTPRINTS[idx], np.array([[c, s, -5.0, -5.0]]))[0]) for c, s in cs)
    assert sum(len(o) + len(a) for o, a in bank) < full


def test_error_codes_and_messages(cb, ctx):
    """Error behaviour at the boundary: invalid arguments come back as COVER_E_INVALID_ARG (-1, the
    reference's std::invalid_argument) with a message, never as a crash; handles are context-bound."""
    with pytest.raises(cb.CoverError) as e:
        cb.Costmap(ctx, np.zeros((10, 10), dtype=np.uint8), 0.0)
    assert e.value.code == -1 and "resolution" in str(e.value).lower()
    with pytest.raises(cb.CoverError):
        cb.Costmap(ctx, np.zeros((10, 10), dtype=np.uint8), -1.0)
    cmap = cb.Costmap(ctx, np.zeros((64, 64), dtype=np.uint8), 0.05)
    th = np.array([[1.0, 0.0]])
    with pytest.raises(cb.CoverError) as e:  # lattice sub-range outside the lattice
        cmap.area_is_free(RECT, cb.Lattice(0.5, 0.05, 10, 0.5, 0.05, 10, th), first=50, count=100)
    assert e.value.code == -1
    with pytest.raises(cb.CoverError):  # more vertices than any kernel supports
        cmap.area_is_free(np.zeros((2, 5000)), np.array([[1.0, 0.0, 1.0, 1.0]]))
    other = cb.Context(0)
    try:
        with pytest.raises(cb.CoverError) as e:  # a cell list of another context
            cb.DiscretePolygon(other, [(0, 0), (1, 0)]).is_free(cmap, [(3, 3)])
        assert e.value.code == -1
    finally:
        other.close()
    with pytest.raises(cb.CoverError):  # footprint index out of range
        cb.FootprintBank(ctx, [(np.zeros((0, 2), np.int32), np.array([[0, 0]], np.int32))]).is_free(cmap, [(1, 1)], which=[3])
    # empty batches are fine
    r = cmap.area_is_free(RECT, np.zeros((0, 4)))
    assert len(r.is_free) == 0
    assert len(cmap.ray_is_free(np.zeros((0, 4), dtype=np.int32)).is_free) == 0
    # the context is still usable after the errors
    assert cmap.area_is_free(RECT, np.array([[1.0, 0.0, 1.5, 1.5]])).is_free[0] == 1


def test_trajectory_scoring(cb, ctx):
    """cover_trajectory_reduce against a numpy restatement: rollouts of 37 poses along arcs."""
    rng = np.random.default_rng(97)
    data = mixed_map(rng, 600, 600, lethal=0.0015)
    cmap = cb.Costmap(ctx, data, 0.05)
    n_traj, steps = 3001, 37
    x0, y0, th0 = rng.uniform(2, 28, n_traj), rng.uniform(2, 28, n_traj), rng.uniform(-math.pi, math.pi, n_traj)
    v, w = rng.uniform(0.05, 0.25, n_traj), rng.uniform(-0.2, 0.2, n_traj)
    k = np.arange(steps)
    th = th0[:, None] + w[:, None] * k
    poses = np.stack([np.cos(th), np.sin(th), x0[:, None] + v[:, None] * k * np.cos(th), y0[:, None] + v[:, None] * k * np.sin(th)], axis=-1).reshape(-1, 4)
    r = cmap.area_accumulate_cost(RECT, poses)
    first, score = cb.trajectory_reduce(ctx, r.is_free, r.cost, steps)
    fr, co = r.is_free.reshape(n_traj, steps), r.cost.reshape(n_traj, steps)
    exp_first = np.where((fr == 0).any(1), (fr == 0).argmax(1), steps)
    neg = ~(co >= 0)
    exp_score = np.where(neg.any(1), co[np.arange(n_traj), neg.argmax(1)], np.where(neg, 0.0, co).sum(1))
    np.testing.assert_array_equal(first, exp_first)
    np.testing.assert_array_equal(score, exp_score)
    assert 0 < (exp_first < steps).mean() < 1
    torch = pytest.importorskip("torch")
    dfirst, dscore = cb.trajectory_reduce(ctx, torch.from_numpy(r.is_free).cuda(), torch.from_numpy(r.cost).cuda(), steps)
    ctx.synchronize()
    np.testing.assert_array_equal(dfirst.cpu().numpy(), exp_first)
    np.testing.assert_array_equal(dscore.cpu().numpy(), exp_score)


def test_gpu_inflation_layer(cb, ctx):
    """cover_map_inflate against a numpy restatement of InflationLayer (min Euclidean cell distance to a lethal
    cell -> cost LUT; NO_INFORMATION stays unless the new cost is >= 253, else max(old, new))."""
    from cover_b200 import workloads as wl
    rng = np.random.default_rng(98)
    sx, sy = 333, 277
    raw = np.zeros((sy, sx), dtype=np.uint8)
    obs = rng.random((sy, sx)) < 6e-4
    raw[obs] = 254
    raw[(rng.random((sy, sx)) < 0.02) & ~obs] = 255        # unknown cells
    raw[(rng.random((sy, sx)) < 0.05) & (raw == 0)] = 120   # some other layer's costs
    raw[0, 0] = raw[sy - 1, sx - 1] = 254                   # obstacles on the border
    rad, lut = wl.inflation_lut(0.05)
    cmap = cb.Costmap(ctx, raw, 0.05)
    cmap.inflate(rad, lut)
    got = cmap.download()
    # numpy restatement
    oy, ox = np.nonzero(raw == 254)
    yy, xx = np.mgrid[0:sy, 0:sx]
    d2 = np.full((sy, sx), 10 ** 9, dtype=np.int64)
    for y, x in zip(oy.tolist(), ox.tolist()):
        y0, y1, x0, x1 = max(0, y - rad), min(sy, y + rad + 1), max(0, x - rad), min(sx, x + rad + 1)
        d2[y0:y1, x0:x1] = np.minimum(d2[y0:y1, x0:x1], (yy[y0:y1, x0:x1] - y) ** 2 + (xx[y0:y1, x0:x1] - x) ** 2)
    cost = np.where(d2 <= rad * rad, lut[np.minimum(d2, rad * rad)], 0).astype(np.uint8)
    reach = d2 <= rad * rad
    expect = raw.copy()
    unknown = raw == 255
    expect[reach & ~unknown] = np.maximum(raw, cost)[reach & ~unknown]
    expect[reach & unknown & (cost >= 253)] = cost[reach & unknown & (cost >= 253)]
    np.testing.assert_array_equal(got, expect)
    # and the inflated map equals the workload generator's stamping formulation on a clean map
    clean = np.zeros((sy, sx), dtype=np.uint8)
    clean[obs] = 254
    c2 = cb.Costmap(ctx, clean, 0.05)
    c2.inflate(rad, lut)
    stamped = np.zeros((sy, sx), dtype=np.uint8)
    yk, xk = np.mgrid[-rad:rad + 1, -rad:rad + 1]
    kern = np.where(yk ** 2 + xk ** 2 <= rad * rad, lut[np.minimum(yk ** 2 + xk ** 2, rad * rad)], 0).astype(np.uint8)
    for y, x in zip(*np.nonzero(obs)):
        y0, y1, x0, x1 = max(0, y - rad), min(sy, y + rad + 1), max(0, x - rad), min(sx, x + rad + 1)
        v = stamped[y0:y1, x0:x1]
        np.maximum(v, kern[y0 - (y - rad):y1 - (y - rad), x0 - (x - rad):x1 - (x - rad)], out=v)
    np.testing.assert_array_equal(c2.download(), stamped)


@pytest.mark.parametrize("res,inflation_radius,density", [(0.05, 1.0, 6e-4), (0.05, 1.0, 0.08), (0.025, 0.8, 2e-3), (0.025, 0.775, 2e-3), (0.02, 1.0, 1e-3), (0.05, 0.05, 0.3)])
def test_gpu_inflation_layer_radii_and_dense_maps(cb, ctx, res, inflation_radius, density):
    """The separable exact distance transform (radius <= 31 cells, in place) and the brute-force kernel (larger radii)
    against scipy's exact Euclidean distance transform: sparse points, walls / dense clutter, radius 1, 20, 32 (brute force) and 50."""
    from scipy import ndimage
    from cover_b200 import workloads as wl
    rng = np.random.default_rng(int(1000 * res) + int(1e4 * density))
    sx, sy = 419, 307
    raw = np.zeros((sy, sx), dtype=np.uint8)
    obs = rng.random((sy, sx)) < density
    obs[100:103, 50:300] = True  # a wall
    obs[0, :] = True             # the map border
    raw[obs] = 254
    raw[(rng.random((sy, sx)) < 0.03) & ~obs] = 255
    raw[(rng.random((sy, sx)) < 0.05) & (raw == 0)] = 97
    rad, lut = wl.inflation_lut(res, inscribed_radius=min(0.3, inflation_radius / 2), inflation_radius=inflation_radius)
    cmap = cb.Costmap(ctx, raw, res)
    cmap.inflate(rad, lut)
    got = cmap.download()
    d2 = np.rint(ndimage.distance_transform_edt(~obs) ** 2).astype(np.int64)
    reach = d2 <= rad * rad
    cost = np.where(reach, lut[np.minimum(d2, rad * rad)], 0).astype(np.uint8)
    expect = raw.copy()
    unknown = raw == 255
    expect[reach & ~unknown] = np.maximum(raw, cost)[reach & ~unknown]
    expect[reach & unknown & (cost >= 253)] = cost[reach & unknown & (cost >= 253)]
    np.testing.assert_array_equal(got, expect)


def test_generator_offset_overloads(cb, ctx):
    """is_free / accumulate_cost(gen, offset, map) (impl/costmap.hpp:29-35,74-81) for the outline and area generators
    of a transformed polygon: cover_poses.cell_offset, against the reference's own offset overloads (oracle/_ref) and
    the restatement - random poses, a lattice (bit-parallel kernels and the value folds), offsets that push cells off
    the map."""
    rng = np.random.default_rng(123)
    data = mixed_map(rng, 500, 420, lethal=0.002)
    cmap = cb.Costmap(ctx, data, 0.05, -1.0, 0.5)
    omap = po.Costmap(data, 0.05, -1.0, 0.5)
    poses = poses_around(rng, 60_000, -1.0, 20.0)
    th = np.concatenate([-math.pi + 2 * math.pi * np.arange(7) / 7, [0.0, math.pi / 2]])
    desc = dict(x0=0.0, dx=0.05, nx=300, y0=1.0, dy=0.05, ny=120, cs=np.stack([np.cos(th), np.sin(th)], 1))
    backends = [po.default()] + ([po.reference()] if po.reference() is not None else [])
    for offset in ((0, 0), (7, -3), (-40, 55), (120, 0)):
        for poly in (RECT, FOOTPRINTS[6], FOOTPRINTS[5]):
            for shape, prefix in ((po.SHAPE_AREA, "area"), (po.SHAPE_OUTLINE, "outline")):
                for op_name, op in (("is_free", po.OP_IS_FREE), ("accumulate_cost", po.OP_ACCUMULATE)):
                    for p_gpu, p_cpu, what in ((poses, poses, "random poses"), (cb.Lattice(**desc), po.Lattice(**desc), "lattice")):
                        gpu = getattr(cmap, f"{prefix}_{op_name}")(poly, p_gpu, offset=offset)
                        for be in backends:
                            r = be.polygon_batch(omap, poly, p_cpu, shape=shape, op=op, threads=8, offset=offset)
                            ok = (gpu.status == 0) & (r.status == 0)
                            np.testing.assert_array_equal(gpu.status == 1, r.status == 1)
                            np.testing.assert_array_equal(gpu.is_free[ok], r.is_free[ok], err_msg=f"{be.name} {prefix} {op_name} {what} offset {offset}")
                            if op == po.OP_ACCUMULATE:
                                np.testing.assert_array_equal(gpu.cost[ok], r.cost[ok], err_msg=f"{be.name} {prefix} {op_name} {what} offset {offset}")
    # an offset is not a shifted origin: cells are truncated towards zero BEFORE the offset is added
    near_zero = np.array([[1.0, 0.0, -0.98, 0.53]])  # vertices land in (-1, 0) map cells before the offset
    a = cmap.area_is_free(RECT * 0.01, near_zero, offset=(3, 3))
    b = po.polygon_batch(omap, RECT * 0.01, near_zero, offset=(3, 3))
    np.testing.assert_array_equal(a.is_free, b.is_free)


def test_area_with_holes(cb, ctx):
    """area_generator(res, polygon_vec): the reference's 'head' with nested holes (test/expand.cpp:241-279) and
    random outline sets, all three ops, against the oracle."""
    rng = np.random.default_rng(99)
    data = mixed_map(rng, 400, 380, lethal=0.002)
    cmap = cb.Costmap(ctx, data, 0.05, -0.3, 0.2)
    omap = po.Costmap(data, 0.05, -0.3, 0.2)
    poses = poses_around(rng, 3000, 0.0, 18.0)
    sets = [head_polygons(), [FOOTPRINTS[0] * 0.5, RECT * 0.5], [RECT, RECT * 0.5, RECT * 0.2],
            [np.array([[0.0, 0.6], [0.1, 0.1]]), RECT], [np.zeros((2, 0)), FOOTPRINTS[6]], [np.array([[0.2], [0.2]]), np.array([[0.25], [0.2]])]]
    names = {"is_free": po.OP_IS_FREE, "accumulate_cost": po.OP_ACCUMULATE, "max_cost": po.OP_MAX}
    for outlines in sets:
        for name, op in names.items():
            gpu = cmap.area_multi(outlines, poses, op=name)
            cpu = po.multi_polygon_batch(omap, outlines, poses, op=op, threads=8)
            assert_same(gpu, cpu, name != "is_free", f"{len(outlines)} outlines {name}")
            ref = po.reference()
            if ref is not None and op != po.OP_MAX and all(o.shape[1] > 0 for o in outlines):  # the reference's own constructor (generators.cpp:133-135)
                r = ref.multi_polygon_batch(omap, outlines, poses, op=op, threads=8)
                ok = gpu.status == 0
                np.testing.assert_array_equal((gpu.status == cb.STATUS_LOGIC_ERROR), (r.status == po.LOGIC_ERROR))
                np.testing.assert_array_equal(gpu.is_free[ok], r.is_free[ok], err_msg=f"vs reference: {len(outlines)} outlines {name}")
                if name != "is_free":
                    np.testing.assert_array_equal(gpu.cost[ok], r.cost[ok])
    # a single outline through the multi entry equals the single-polygon entry
    a = cmap.area_multi([FOOTPRINTS[3] * 0.5], poses, op="accumulate_cost")
    b = cmap.area_accumulate_cost(FOOTPRINTS[3] * 0.5, poses)
    np.testing.assert_array_equal(a.cost, b.cost)
    none = cmap.area_multi(head_polygons(), None)
    assert_same(none, po.multi_polygon_batch(omap, head_polygons(), None))


def test_plain_c_client(cb, tmp_path):
    """tests/c/abi_smoke.c through gcc -std=c99: hand-computable answers via the raw C ABI."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.dirname(cb.library_path())
    exe = tmp_path / "abi_smoke"
    build = subprocess.run(["gcc", "-std=c99", "-I", os.path.join(root, "include"), os.path.join(root, "tests", "c", "abi_smoke.c"), "-o", str(exe),
                            "-L", libdir, "-lcover_b200", f"-Wl,-rpath,{libdir}", "-lm"], capture_output=True, text=True)
    assert build.returncode == 0, build.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert run.returncode == 0, run.stdout + run.stderr
