"""Differential fuzz: the CPU restatement (oracle/cover_oracle.hpp) against the REFERENCE'S OWN sources
compiled over shim headers (oracle/_ref/libcover_ref.so, built by oracle/Makefile when /root/reference
is present).  Skipped when oracle/_ref was not built.  Sequences are compared element by element, so
the generator ORDER and duplicate cells are pinned, not just the cell sets.
"""
import math

import numpy as np
import pytest

from oracle import pyoracle as po
from tests.fixtures import FOOTPRINTS, RECT, head_polygons

ref = po.reference()
pytestmark = pytest.mark.skipif(ref is None, reason="oracle/_ref not built (reference tree absent)")
port = po.default()


def random_map(rng, sx, sy, lethal=0.01):
    data = rng.integers(0, 253, (sy, sx), dtype=np.uint8)
    r = rng.random((sy, sx))
    data[r < lethal] = 254
    data[(r >= lethal) & (r < 1.5 * lethal)] = 255
    data[(r >= 1.5 * lethal) & (r < 2.5 * lethal)] = 253
    return data


def random_poses(rng, n, lo, hi):
    th = rng.uniform(-math.pi, math.pi, n)
    return np.stack([np.cos(th), np.sin(th), rng.uniform(lo, hi, n), rng.uniform(lo, hi, n)], axis=1)


def test_rays_and_outlines_random():
    rng = np.random.default_rng(1)
    for _ in range(500):
        b, e = rng.integers(-40, 40, 2), rng.integers(-40, 40, 2)
        assert port.ray_cells(b, e).tolist() == ref.ray_cells(b, e).tolist()
    for _ in range(500):
        n = int(rng.integers(1, 9))
        sparse = rng.integers(0, 7, (n, 2)).astype(np.int32)
        assert port.outline_cells(sparse).tolist() == ref.outline_cells(sparse).tolist()


def test_area_random_small_polygons_including_throws():
    """Small grids provoke every quirk: cusps, ties, degenerate edges, open outlines (logic_error)."""
    rng = np.random.default_rng(2)
    threw = 0
    for _ in range(4000):
        n = int(rng.integers(1, 8))
        sparse = rng.integers(0, 6, (n, 2)).astype(np.int32)
        dense = port.outline_cells(sparse)
        try:
            a = port.area_cells([dense]).tolist()
        except RuntimeError:
            a = "throw"
            threw += 1
        try:
            b = ref.area_cells([dense]).tolist()
        except RuntimeError:
            b = "throw"
        assert a == b, sparse.tolist()
    assert threw > 0


def test_area_random_larger_polygons():
    rng = np.random.default_rng(3)
    for _ in range(600):
        n = int(rng.integers(3, 24))
        sparse = rng.integers(-30, 30, (n, 2)).astype(np.int32)
        dense = port.outline_cells(sparse)
        try:
            a = port.area_cells([dense]).tolist()
        except RuntimeError:
            a = "throw"
        try:
            b = ref.area_cells([dense]).tolist()
        except RuntimeError:
            b = "throw"
        assert a == b


def test_area_multi_outline_holes():
    for res in (0.2, 0.1, 0.05):
        denses = [port.outline_cells(port.to_discrete(res, p)) for p in head_polygons()]
        assert port.area_cells(denses).tolist() == ref.area_cells(denses).tolist()


def test_area_multi_outline_checks_against_the_reference():
    """is_free / accumulate_cost over area_generator(res, polygon_vec) (generators.cpp:126-135) - the reference's own
    constructor through oracle/_ref - against the restatement: the 'head' with nested holes under random poses."""
    rng = np.random.default_rng(77)
    data = random_map(rng, 300, 280, lethal=0.0001)
    cmap = po.Costmap(data, 0.05, -0.5, 0.25)
    poses = random_poses(rng, 400, 0.0, 12.0)
    for op in (po.OP_IS_FREE, po.OP_ACCUMULATE):
        a = port.multi_polygon_batch(cmap, head_polygons(), poses, op=op)
        b = ref.multi_polygon_batch(cmap, head_polygons(), poses, op=op)
        assert a.is_free.tolist() == b.is_free.tolist()
        assert (a.status == po.LOGIC_ERROR).tolist() == (b.status == po.LOGIC_ERROR).tolist()
        if op == po.OP_ACCUMULATE:
            np.testing.assert_array_equal(a.cost, b.cost)
        else:
            assert 0.02 < a.is_free.mean() < 0.98
    a = port.multi_polygon_batch(cmap, head_polygons(), None)
    b = ref.multi_polygon_batch(cmap, head_polygons(), None)
    assert a.is_free.tolist() == b.is_free.tolist()


@pytest.mark.parametrize("idx", range(7))
@pytest.mark.parametrize("shape", [po.SHAPE_AREA, po.SHAPE_OUTLINE])
def test_polygon_pose_checks(idx, shape):
    """is_free(polygon, pose, map) (costmap.cpp:78-85) and accumulate_cost over the generators."""
    rng = np.random.default_rng(100 + idx)
    data = random_map(rng, 400, 300, lethal=0.0005)
    cmap = po.Costmap(data, 0.05, -1.0, -2.0)
    poses = random_poses(rng, 300, -2.0, 16.0)  # some poses stick out of the map
    for op in (po.OP_IS_FREE, po.OP_ACCUMULATE):
        a = port.polygon_batch(cmap, FOOTPRINTS[idx], poses, shape=shape, op=op)
        b = ref.polygon_batch(cmap, FOOTPRINTS[idx], poses, shape=shape, op=op)
        assert a.is_free.tolist() == b.is_free.tolist()
        assert (a.status == po.LOGIC_ERROR).tolist() == (b.status == po.LOGIC_ERROR).tolist()
        if op == po.OP_ACCUMULATE:
            np.testing.assert_array_equal(a.cost, b.cost)
    a = port.polygon_batch(cmap, FOOTPRINTS[idx] * 0.2 + 5.0, None, shape=shape)
    b = ref.polygon_batch(cmap, FOOTPRINTS[idx] * 0.2 + 5.0, None, shape=shape)
    assert a.is_free.tolist() == b.is_free.tolist()


def test_rect_lattice_checks():
    rng = np.random.default_rng(5)
    cmap = po.Costmap(random_map(rng, 300, 300, lethal=0.002), 0.05, 0.0, 0.0)
    th = np.linspace(-math.pi, math.pi, 7, endpoint=False)
    lat = po.Lattice(1.0, 0.05, 40, 1.2, 0.05, 30, np.stack([np.cos(th), np.sin(th)], 1))
    a = port.polygon_batch(cmap, RECT, lat)
    b = ref.polygon_batch(cmap, RECT, lat)
    assert a.is_free.tolist() == b.is_free.tolist()
    assert 0 < a.is_free.mean() < 1


def test_rays_cells_and_offsets():
    rng = np.random.default_rng(6)
    cmap = po.Costmap(random_map(rng, 64, 48, lethal=0.02), 1.0)
    segs = rng.integers(-5, 70, (2000, 4)).astype(np.int32)
    for op in (po.OP_IS_FREE, po.OP_ACCUMULATE):
        for off in (None, (3, -2)):
            a, b = port.rays_batch(cmap, segs, off, op), ref.rays_batch(cmap, segs, off, op)
            assert a.is_free.tolist() == b.is_free.tolist()
            np.testing.assert_array_equal(a.cost, b.cost)
    cells = port.area_cells([port.outline_cells(port.to_discrete(0.2, FOOTPRINTS[5]))])
    offs = rng.integers(-5, 60, (2000, 2)).astype(np.int32)
    assert port.cells_batch(cmap, cells, offs).is_free.tolist() == ref.cells_batch(cmap, cells, offs).is_free.tolist()


def test_split_footprints():
    """continuous_footprint::is_free (footprints.cpp:130-196) and discrete_footprint (198-247) over a
    hand-made split (the boost split itself is out of the path)."""
    rng = np.random.default_rng(8)
    data = random_map(rng, 300, 300, lethal=0.001)
    cmap = po.Costmap(data, 0.05, -2.0, -1.0)
    toru = FOOTPRINTS[5]
    outlines = [toru * 0.4, toru * 0.2 + np.array([[0.3], [0.0]])]
    areas = [toru, RECT * 0.5 + np.array([[-0.2], [0.1]])]
    poses = random_poses(rng, 1500, -2.5, 13.5)
    a = port.continuous_footprint_batch(cmap, outlines, areas, poses)
    b = ref.continuous_footprint_batch(cmap, outlines, areas, poses)
    assert a.is_free.tolist() == b.is_free.tolist()
    assert 0 < a.is_free.mean() < 1

    bank = []
    for k in range(6):
        th = 2 * math.pi * k / 6
        R = np.array([[math.cos(th), -math.sin(th)], [math.sin(th), math.cos(th)]])
        oc_a, ac_a = port.make_discrete_footprint(0.05, [R @ o for o in outlines], [R @ p for p in areas])
        oc_b, ac_b = ref.make_discrete_footprint(0.05, [R @ o for o in outlines], [R @ p for p in areas])
        assert oc_a.tolist() == oc_b.tolist() and ac_a.tolist() == ac_b.tolist()
        bank.append((oc_a, ac_a))
    offs = rng.integers(-10, 310, (3000, 2)).astype(np.int32)
    which = rng.integers(0, 6, 3000).astype(np.int32)
    a = port.discrete_footprint_batch(cmap, bank, offs, which)
    b = ref.discrete_footprint_batch(cmap, bank, offs, which)
    assert a.is_free.tolist() == b.is_free.tolist()
    assert 0 < a.is_free.mean() < 1
