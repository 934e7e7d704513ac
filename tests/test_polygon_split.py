"""get_radius / continuous_footprint(radius, polygon) without boost::geometry (cover_b200/csrc/polygon_split.hpp through
the C ABI): the reference's property tests test/footprints.cpp:20-166 (host only, no GPU), and - on the GPU - its
safety regression :168-242 with the GPU inflation layer standing in for costmap_2d::InflationLayer.

Deviation from the reference's `circle` test (:150-166), on purpose: boost's difference returns the 16-gon with the
band as an INNER ring there (the band is tangent to all edges), which from_boost drops - so the reference ends up with
no area at all and never looks at the 16 corner tips beyond the band.  This implementation returns those tips as areas.
"""
import math

import numpy as np
import pytest

from tests.fixtures import FOOTPRINTS


@pytest.fixture(scope="module")
def cb():
    import cover_b200
    return cover_b200


def contains(ring, pts):
    """Even-odd point-in-polygon of pts [N, 2] for a ring [2, V] (open or closed)."""
    x, y = ring[0], ring[1]
    if x[0] != x[-1] or y[0] != y[-1]:
        x, y = np.append(x, x[0]), np.append(y, y[0])
    inside = np.zeros(len(pts), dtype=bool)
    for i in range(len(x) - 1):
        crosses = (y[i] > pts[:, 1]) != (y[i + 1] > pts[:, 1])
        with np.errstate(divide="ignore", invalid="ignore"):
            xi = x[i] + (pts[:, 1] - y[i]) * (x[i + 1] - x[i]) / (y[i + 1] - y[i])
        inside ^= crosses & (pts[:, 0] < xi)
    return inside


def area_of(ring):
    return 0.5 * abs(np.sum(ring[0, :-1] * ring[1, 1:] - ring[0, 1:] * ring[1, :-1]))


def test_get_radius_empty_and_box(cb):
    assert cb.get_radius(np.zeros((2, 0))) == 0  # test/footprints.cpp:20
    p = np.array([[-1.0, -1, 1, 1], [-2, 2, 2, -2]])
    assert cb.get_radius(p) == pytest.approx(1.0, rel=1e-6)
    for ii in range(100):  # invariant to rigid transforms (test/footprints.cpp:31-38)
        rad = ii / 50.0 * math.pi
        R = np.array([[math.cos(rad), -math.sin(rad)], [math.sin(rad), math.cos(rad)]])
        assert cb.get_radius(R @ p + np.array([[ii], [-ii]])) == pytest.approx(1.0, rel=1e-6)


def test_degenerate_inputs_give_an_empty_split(cb):
    assert cb.split_polygon(np.zeros((2, 0)), 1.0) == ([], [])  # an empty polygon is a no-op (:42-50)
    rng = np.random.default_rng(1)
    for cols in (1, 2):  # fewer than three distinct points (:60-68)
        assert cb.split_polygon(rng.uniform(-1, 1, (2, cols)), 1.0) == ([], [])
    assert cb.split_polygon(np.zeros((2, 10)), 1.0) == ([], [])  # all points the same (:70-77)


def test_invalid_radius_throws(cb):
    with pytest.raises(ValueError):  # std::invalid_argument("Radius must be positive") (:79-83)
        cb.split_polygon(np.zeros((2, 2)), -1.0)
    with pytest.raises(ValueError):
        cb.split_polygon(np.array([[0.0, 1, 0], [0, 0, 1]]), 0.0)


def test_too_large_radius_returns_the_polygon_as_the_one_area(cb):
    p = np.array([[0.0, 0.5, -0.5, 0.0], [1.0, 0.0, 0.0, 1.0]])  # (:85-100) closed and clockwise, like boost's output
    outlines, areas = cb.split_polygon(p, 2.0)
    assert outlines == [] and len(areas) == 1
    assert np.abs(p - areas[0]).sum() == 0.0


def test_three_points_are_enough(cb):
    rng = np.random.default_rng(2)
    for _ in range(20):  # (:102-109)
        outlines, areas = cb.split_polygon(rng.uniform(-1, 1, (2, 3)))
        assert outlines and areas


def test_triangle_rectangle_circle_outline_sits_at_the_incentre(cb):
    r = 1.0
    a, h = r / math.tan(math.radians(30)), r / math.sin(math.radians(30))
    outlines, areas = cb.split_polygon(np.array([[0.0, 2 * a, a], [0.0, 0.0, h + r]]))  # (:113-131)
    assert outlines and areas
    assert outlines[0][:, :-1].mean(axis=1) == pytest.approx([a, r], abs=1e-2)
    outlines, areas = cb.split_polygon(np.array([[0.0, 2, 2, 0], [0.0, 0, 2, 2]]))  # (:133-148)
    assert outlines and areas
    assert outlines[0][:, :-1].mean(axis=1) == pytest.approx([1.0, 1.0], abs=1e-2)
    ang = 2 * math.pi * np.arange(16) / 16
    circle = np.stack([np.cos(ang), np.sin(ang)])
    outlines, areas = cb.split_polygon(circle)  # (:150-166; see the module docstring for the areas)
    assert outlines
    assert outlines[0][:, :-1].mean(axis=1) == pytest.approx([0.0, 0.0], abs=1e-2)
    for ring in areas:  # the tips of the 16-gon beyond the band, nothing else
        assert np.hypot(*ring).min() >= math.cos(math.pi / 16) - 2e-2 and area_of(ring) < 5e-3


def test_split_covers_the_polygon_and_outlines_are_inscribed(cb):
    """The construction's two properties, checked on a point grid for the reference's footprints: (safety) every point of
    P lies in an area or within `radius` + h of the first outline ring (h = the grid pitch of the construction, 1/384 of
    the footprint's extent: where the band is tangent to P's boundary the area's tail gets thinner than the grid
    resolves); (no false alarm) outline rings keep >= radius - 3 h from P's boundary."""
    for idx in (0, 4, 5, 6):
        poly = FOOTPRINTS[idx]
        radius = cb.get_radius(poly)
        outlines, areas = cb.split_polygon(poly)
        assert outlines
        lo, hi = poly.min(axis=1), poly.max(axis=1)
        h = max(hi - lo) / 384.0
        xs, ys = np.meshgrid(np.linspace(lo[0], hi[0], 160), np.linspace(lo[1], hi[1], 160))
        pts = np.stack([xs.ravel(), ys.ravel()], 1)
        inside = contains(poly, pts)
        ring = outlines[0].T

        def dist_to_ring(ring, pts):
            a, b = ring[:-1], ring[1:]
            v, w = b - a, pts[:, None, :] - a[None]
            t = np.clip((w * v).sum(-1) / np.maximum((v * v).sum(-1), 1e-300), 0, 1)
            return np.linalg.norm(w - t[..., None] * v, axis=-1).min(axis=1)
        near = dist_to_ring(ring, pts) <= radius + h
        in_area = np.zeros(len(pts), dtype=bool)
        for a in areas:
            in_area |= contains(a, pts) | (dist_to_ring(a.T, pts) <= h)  # (h of slack: the rings are polylines of a curved boundary)
        assert (inside & ~near & ~in_area).sum() == 0, f"footprint {idx}: points neither near the outline nor in an area"
        closed = np.concatenate([poly, poly[:, :1]], axis=1).T
        for o in outlines:
            assert dist_to_ring(closed, o.T).min() >= radius * (1 - 1e-6) - 3 * h


@pytest.mark.gpu
@pytest.mark.parametrize("idx", [4, 5])
def test_safety_regression_with_the_inflation_layer(cb, idx):
    """test/footprints.cpp:168-242: every cell inside the footprint, marked lethal and inflated (GPU inflation layer,
    inscribed radius = the footprint's, like costmap_2d's setFootprint), must make the split footprint collide at the
    identity pose; the reference tolerates 2 misses "due to discretization issues".  With `slack` = one cell this
    construction meets that bar; the plain construction (slack 0: areas cut where their tails along P's edges get thinner
    than the construction grid, no allowance for the cell metric) is held to 1 % of the cells."""
    from oracle import pyoracle as po
    from cover_b200 import workloads as wl
    poly = FOOTPRINTS[idx]
    res = 0.05
    closed = np.concatenate([poly, poly[:, :1]], axis=1)
    a, b = closed[:, :-1].T, closed[:, 1:].T
    t = np.clip((-a * (b - a)).sum(1) / np.maximum(((b - a) ** 2).sum(1), 1e-300), 0, 1)
    inscribed = np.linalg.norm(a + t[:, None] * (b - a), axis=1).min()  # min distance from the robot origin to the edges
    rad, lut = wl.inflation_lut(res, inscribed_radius=inscribed, inflation_radius=1.0)
    ctx = cb.Context(0)
    cells = po.area_cells([po.outline_cells(po.to_discrete(res, poly + 5.0))])  # the footprint's cells on the 200 x 200 map at (-5, -5)
    pose = np.array([[1.0, 0.0, 0.0, 0.0]])
    empty = np.zeros((200, 200), dtype=np.uint8)
    cmap = cb.Costmap(ctx, empty, res, -5.0, -5.0)
    for slack, allowed in ((res, 2), (0.0, max(2, len(cells) // 100))):
        outlines, areas = cb.split_polygon(poly, slack=slack)
        assert outlines and areas
        sf = cb.SplitFootprint(ctx, outlines, areas)
        cmap.upload(empty)
        assert sf.is_free(cmap, pose).is_free[0] == 1
        misses = 0
        for x, y in cells.tolist():
            data = empty.copy()
            data[y, x] = 254
            cmap.upload(data)
            cmap.inflate(rad, lut)
            misses += int(sf.is_free(cmap, pose).is_free[0])
        assert misses <= allowed, f"slack {slack}: {misses} of {len(cells)} single-obstacle maps were missed"
    ctx.close()
