"""Pins the CPU oracle against the reference's own known-answer tests.

Each test cites the reference test it restates (paths relative to /root/reference).  Comparators that
are ROS code in the reference (FootprintHelper::getLineCells, polygonOutlineCells, convexFillCells) are
replaced by an independent textbook Bresenham / cv2.fillPoly; the OpenCV comparisons are the
reference's own (test/expand.cpp:179-317).
"""
import math

import numpy as np
import pytest

from oracle import pyoracle as po
from tests.fixtures import FOOTPRINTS, FOOTPRINT_NAMES, SIMPLE_FOOTPRINTS, head_polygons

LETHAL, NO_INFO, INSCRIBED = 254, 255, 253


def textbook_line(x0, y0, x1, y1):
    """Classic integer Bresenham as used by ROS base_local_planner::FootprintHelper::getLineCells
    (independent comparator, end point included)."""
    dx, dy = abs(x1 - x0), abs(y1 - y0)
    x, y = x0, y0
    sx = 1 if x1 >= x0 else -1
    sy = 1 if y1 >= y0 else -1
    xinc1 = xinc2 = sx
    yinc1 = yinc2 = sy
    if dx >= dy:
        xinc1 = 0
        yinc2 = 0
        den, num, numadd, n = dx, dx // 2, dy, dx
    else:
        xinc2 = 0
        yinc1 = 0
        den, num, numadd, n = dy, dy // 2, dx, dy
    pts = []
    for _ in range(n + 1):
        pts.append((x, y))
        num += numadd
        if num >= den:
            num -= den
            x += xinc1
            y += yinc1
        x += xinc2
        y += yinc2
    return pts


# ---- ray_generator: test/generators.cpp:21-83 ------------------------------------------------

def test_ray_empty():
    assert len(po.ray_cells((0, 0), (0, 0))) == 0


def test_ray_simple():
    cells = po.ray_cells((-5, 0), (5, 0))
    assert len(cells) == 10
    assert cells[:, 0].tolist() == list(range(-5, 5)) and not cells[:, 1].any()


@pytest.mark.parametrize("seg", [(0, 0, 100, 30), (0, 0, 100, 100), (30, 100, 0, 0), (0, 100, 100, 30), (10, -11, -10, 30),
                                 (0, 0, -10, -30), (-10, -30, 100, 30), (10, -30, -100, 30), (0, 2, 2, 2)])
def test_ray_matches_textbook_bresenham(seg):
    cells = po.ray_cells(seg[:2], seg[2:])
    expected = textbook_line(*seg)
    assert len(cells) == len(expected) - 1  # end excluded
    assert [tuple(c) for c in cells.tolist()] == expected[:-1]


def test_ray_random_vs_textbook():
    rng = np.random.default_rng(7)
    for _ in range(2000):
        x0, y0, x1, y1 = rng.integers(-60, 60, 4).tolist()
        cells = po.ray_cells((x0, y0), (x1, y1))
        assert [tuple(c) for c in cells.tolist()] == textbook_line(x0, y0, x1, y1)[:-1]


# ---- outline_generator: test/generators.cpp:89-179 -------------------------------------------

def test_outline_empty():
    assert len(po.outline_cells(np.zeros((0, 2), dtype=np.int32))) == 0


def test_outline_one_point():
    cells = po.outline_cells(np.ones((10, 2), dtype=np.int32))
    assert cells.tolist() == [[1, 1]]


def test_outline_two_points():
    sparse = np.ones((10, 2), dtype=np.int32)
    sparse[9] = (2, 3)
    ray = po.ray_cells((1, 1), (2, 3))
    cells = po.outline_cells(sparse)
    assert len(cells) == len(ray) + 1
    assert cells[:-1].tolist() == ray.tolist()
    assert cells[-1].tolist() == [2, 3]


def test_outline_box_order():
    rows, cols = 10, 20
    cells = po.outline_cells([(0, 0), (rows, 0), (rows, cols), (0, cols)])
    assert len(cells) == (rows + cols) * 2
    exp = [(i, 0) for i in range(rows)] + [(rows, i) for i in range(cols)]
    exp += [(rows - i, cols) for i in range(rows)] + [(0, cols - i) for i in range(cols)]
    assert [tuple(c) for c in cells.tolist()] == exp


@pytest.mark.parametrize("idx", range(7))
def test_outline_set_matches_textbook_edges(idx):
    """test/generators.cpp:181-215 (comparator polygonOutlineCells = Bresenham over every edge)."""
    fp = FOOTPRINTS[idx].copy()
    fp -= fp.min(axis=1, keepdims=True)
    fp += 0.5
    sparse = po.to_discrete(0.05, fp)
    got = {tuple(c) for c in po.outline_cells(sparse).tolist()}
    exp = set()
    for i in range(len(sparse)):
        a, b = sparse[i], sparse[(i + 1) % len(sparse)]
        exp.update(textbook_line(int(a[0]), int(a[1]), int(b[0]), int(b[1])))
    assert got == exp


# ---- area_generator: test/generators.cpp:221-268 ---------------------------------------------

def test_area_empty():
    assert len(po.area_cells([])) == 0
    assert len(po.area_cells([np.zeros((0, 2), dtype=np.int32)])) == 0


def test_area_box_row_major():
    rows, cols = 10, 20
    dense = po.outline_cells([(0, 0), (rows, 0), (rows, cols), (0, cols)])
    cells = po.area_cells([dense])
    assert len(cells) == (rows + 1) * (cols + 1)
    idx = cells[:, 0] + cells[:, 1] * (rows + 1)
    assert idx.tolist() == list(range(len(cells)))


def test_area_open_outline_throws():
    """SURVEY F6 / generators.cpp:246-249."""
    with pytest.raises(RuntimeError):
        po.area_cells([po.outline_cells([(0, 0), (2, 3)])])
    with pytest.raises(RuntimeError):
        po.area_cells([po.outline_cells([(0, 0), (0, 5)])])
    assert len(po.area_cells([po.outline_cells([(1, 1)])])) == 1
    assert len(po.area_cells([po.outline_cells([(0, 0), (5, 0)])])) == 6


def test_to_discrete_truncates_toward_zero_and_rejects_bad_resolution():
    """src/cover/base.hpp:50-57."""
    d = po.to_discrete(0.05, np.array([[0.049, -0.049, 0.051, -0.051], [1.0, -1.0, 0.0, -0.0]]))
    assert d.tolist() == [[0, 20], [0, -20], [1, 0], [-1, 0]]
    with pytest.raises(ValueError):
        po.to_discrete(0.0, np.zeros((2, 1)))
    with pytest.raises(ValueError):
        po.to_discrete(-1.0, np.zeros((2, 1)))


# ---- fill vs OpenCV: test/expand.cpp:116-317 -------------------------------------------------

def _fillpoly_set(outlines, rows, cols):
    cv2 = pytest.importorskip("cv2")
    img = np.zeros((rows, cols), dtype=np.uint8)
    cv2.fillPoly(img, [np.asarray(o, dtype=np.int32).reshape(-1, 1, 2) for o in outlines], 1)
    ys, xs = np.nonzero(img)
    return set(zip(xs.tolist(), ys.tolist()))


def _check_fill(poly, res):
    dense = po.outline_cells(po.to_discrete(res, poly))
    dense = dense - dense.min(axis=0, keepdims=True)
    cells = po.area_cells([dense])
    rows, cols = int(cells[:, 1].max()) + 1, int(cells[:, 0].max()) + 1
    assert {tuple(c) for c in cells.tolist()} == _fillpoly_set([dense], rows, cols)


@pytest.mark.parametrize("res", [0.2, 0.1, 0.05, 0.025])
@pytest.mark.parametrize("idx", range(7))
def test_area_matches_fillpoly_nonconvex(idx, res):
    """test/expand.cpp:179-223 — 7 footprints x 4 resolutions."""
    _check_fill(FOOTPRINTS[idx], res)


@pytest.mark.parametrize("res", [0.1, 0.05, 0.025, 0.01])
@pytest.mark.parametrize("idx", range(5))
def test_area_matches_fill_convex_inputs(idx, res):
    """test/expand.cpp:116-172 inputs (comparator convexFillCells replaced by fillPoly)."""
    _check_fill(SIMPLE_FOOTPRINTS[idx], res)


@pytest.mark.parametrize("res", [0.2, 0.1, 0.05, 0.025, 0.01])
def test_area_with_holes_matches_fillpoly(res):
    """test/expand.cpp:225-317 — 7 nested outlines, generalized even-odd rule."""
    denses = [po.outline_cells(po.to_discrete(res, p)) for p in head_polygons()]
    cells = po.area_cells(denses)
    rows, cols = int(cells[:, 1].max()) + 1, int(cells[:, 0].max()) + 1
    assert {tuple(c) for c in cells.tolist()} == _fillpoly_set(denses, rows, cols)


# ---- costmap functors and checks: test/costmap.cpp:35-194 ------------------------------------

def _fixture_map():
    data = np.zeros((20, 10), dtype=np.uint8)  # Costmap2D(10, 20, 1, 0, 0): size_x=10, size_y=20
    data[5, 5] = LETHAL
    data[6, 5] = NO_INFO
    data[7, 5] = INSCRIBED
    return po.Costmap(data, 1.0, 0.0, 0.0)


def _single(cmap, cells, offset=(0, 0), op=po.OP_IS_FREE):
    r = po.cells_batch(cmap, cells, [offset], op=op)
    return bool(r.is_free[0]), float(r.cost[0])


def test_is_inside_and_lethal_truth_table():
    m = _fixture_map()
    # is_inside corners (test/costmap.cpp:35-51) through is_lethal: outside => lethal
    for cell, free in [((0, 0), True), ((9, 19), True), ((10, 19), False), ((9, 20), False), ((0, -1), False), ((-1, 0), False)]:
        assert _single(m, [cell])[0] is free
    # is_lethal (test/costmap.cpp:67-78)
    assert _single(m, [(5, 5)])[0] is False
    assert _single(m, [(5, 6)])[0] is True   # NO_INFORMATION is ok
    assert _single(m, [(5, 7)])[0] is True   # INSCRIBED is ok for is_lethal
    # costmap_model_like_getter (src/cover/costmap.cpp:35-46)
    assert _single(m, [(5, 5)], op=po.OP_ACCUMULATE)[1] == -1
    assert _single(m, [(5, 6)], op=po.OP_ACCUMULATE)[1] == -2
    assert _single(m, [(10, 19)], op=po.OP_ACCUMULATE)[1] == -3
    assert _single(m, [(5, 7)], op=po.OP_ACCUMULATE)[1] == 253
    # get_cell_cost fold (extension)
    assert _single(m, [(0, 0), (5, 7)], op=po.OP_MAX)[1] == 253
    assert _single(m, [(0, 0), (5, 6)], op=po.OP_MAX)[1] == 255


def test_is_lethal_or_inflated_via_footprint_outline():
    """test/costmap.cpp:80-92 through discrete_footprint's outline rule (footprints.cpp:244)."""
    m = _fixture_map()
    none = np.zeros((0, 2), dtype=np.int32)
    for cell, free in [((5, 5), False), ((5, 7), False), ((10, 19), False), ((0, 0), True), ((5, 6), True)]:
        r = po.discrete_footprint_batch(m, [(np.array([cell], dtype=np.int32), none)], [(0, 0)])
        assert bool(r.is_free[0]) is free


def test_is_free_ray():
    """test/costmap.cpp:106-136."""
    m = _fixture_map()
    assert po.rays_batch(m, [(4, 0, 4, 9)]).is_free[0] == 1
    assert po.rays_batch(m, [(5, 0, 5, 9)]).is_free[0] == 0
    assert po.rays_batch(m, [(-4, -4, -4, 4)]).is_free[0] == 0
    assert po.rays_batch(m, [(-4, -4, -4, 4)], offset=(4, 4)).is_free[0] == 1


def test_accumulate_cost():
    """test/costmap.cpp:138-194 (sparse.col(3) is (0, 19): getSizeInMetersY()=19.5 narrowed to int)."""
    m = po.Costmap(np.full((20, 10), 10, dtype=np.uint8), 1.0)
    sparse = np.array([(0, 0), (9, 0), (9, 19), (0, 19)], dtype=np.int32)
    area = po.area_cells([po.outline_cells(sparse)])
    assert _single(m, area, op=po.OP_ACCUMULATE)[1] == 10 * 200
    shifted = po.area_cells([po.outline_cells(sparse - np.array([4, 4], dtype=np.int32))])
    assert _single(m, shifted, op=po.OP_ACCUMULATE)[1] == -3
    assert _single(m, shifted, offset=(4, 4), op=po.OP_ACCUMULATE)[1] == 2000
    assert po.rays_batch(m, [(-1, -1, 5, 5)], op=po.OP_ACCUMULATE).cost[0] == -3


# ---- continuous_footprint::is_free: test/footprints.cpp:244-305 -------------------------------

def _box():
    box = np.array([[0, 1, 1, 0], [0, 0, 2, 2]], dtype=np.float64)
    return box - np.array([[0.5], [1.0]])


def _identity():
    return np.array([[1.0, 0.0, 0.0, 0.0]])


def test_continuous_footprint_outline_vs_area_semantics():
    data = np.zeros((200, 100), dtype=np.uint8)  # Costmap2D(100, 200, 0.1, -5, -10)
    m = po.Costmap(data, 0.1, -5.0, -10.0)
    box = _box()
    assert po.continuous_footprint_batch(m, [box], [], _identity()).is_free[0] == 1
    data[100, 50] = LETHAL  # interior cell: invisible to an outline-only footprint
    assert po.continuous_footprint_batch(m, [box], [], _identity()).is_free[0] == 1
    # worldToMap(box(0,0), box(1,0)) = ((-0.5+5)/0.1, (-1+10)/0.1) = (45, 90)
    data[90, 45] = INSCRIBED
    assert po.continuous_footprint_batch(m, [box], [], _identity()).is_free[0] == 0
    assert po.continuous_footprint_batch(m, [], [box], _identity()).is_free[0] == 0  # lethal inside the area
    data[100, 50] = INSCRIBED
    assert po.continuous_footprint_batch(m, [], [box], _identity()).is_free[0] == 1  # 253 does not stop an area


def test_continuous_footprint_rotation_and_out_of_bounds():
    """test/footprints.cpp:285-305 with a hand-made split (the boost split is out of the path)."""
    m = po.Costmap(np.zeros((200, 100), dtype=np.uint8), 0.1, -5.0, -10.0)
    box = _box()
    inner = box * 0.5
    for ii in range(10):
        ang = (ii // 5) * math.pi  # the reference's `ii / 5 * M_PI` is integer division
        pose = np.array([[math.cos(ang), math.sin(ang), 0.0, 0.0]])
        assert po.continuous_footprint_batch(m, [inner], [box], pose).is_free[0] == 1
    pose = np.array([[1.0, 0.0, 5.0, 10.0]])  # getSizeInMeters / 2
    assert po.continuous_footprint_batch(m, [inner], [box], pose).is_free[0] == 0


# ---- F5 quirk census: the structural facts the CUDA kernels rely on ---------------------------

def test_rows_never_exceed_16_ranges_on_reference_footprints():
    rng = np.random.default_rng(11)
    for fp in FOOTPRINTS:
        for _ in range(40):
            th = rng.uniform(-math.pi, math.pi)
            R = np.array([[math.cos(th), -math.sin(th)], [math.sin(th), math.cos(th)]])
            q = R @ fp + rng.uniform(20, 30, (2, 1))
            _, mr = po.area_cells([po.outline_cells(po.to_discrete(0.05, q))], return_max_ranges=True)
            assert mr <= 16
