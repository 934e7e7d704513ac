"""cover_b200.split.split_cells on the CPU: the footprint cell set comes from the oracle's to_area here (the
GPU test feeds it from cover_area_expand) and the split is checked with the oracle's discrete_footprint."""
import math

import numpy as np
import pytest

from cover_b200 import split
from oracle import pyoracle as po
from tests.fixtures import FOOTPRINTS, RECT


def footprint_cells(poly, theta, res):
    R = np.array([[math.cos(theta), -math.sin(theta)], [math.sin(theta), math.cos(theta)]])
    return po.area_cells([po.outline_cells(po.to_discrete(res, R @ poly))])


def test_split_partitions_and_shrinks_the_work():
    cells = footprint_cells(FOOTPRINTS[4], 0.4, 0.05)
    outline, area = split.split_cells(cells, 0.05, 0.2)
    fset = {tuple(c) for c in cells.tolist()}
    assert {tuple(c) for c in outline.tolist()} <= fset and {tuple(c) for c in area.tolist()} <= fset
    assert not ({tuple(c) for c in outline.tolist()} & {tuple(c) for c in area.tolist()})
    assert 0 < len(outline) and len(outline) + len(area) < len(fset)


def test_too_large_radius_returns_the_footprint_as_area():
    """test/footprints.cpp:85-101 (continuous_footprint, too_large_radius), grid version."""
    cells = footprint_cells(RECT, 0.2, 0.05)
    outline, area = split.split_cells(cells, 0.05, 2.0)
    assert len(outline) == 0 and len(area) == len(np.unique(cells, axis=0))


def test_invalid_radius_and_empty():
    with pytest.raises(ValueError):
        split.split_cells(footprint_cells(RECT, 0.0, 0.05), 0.05, -1.0)
    o, a = split.split_cells(np.zeros((0, 2), np.int32), 0.05, 1.0)
    assert len(o) == 0 and len(a) == 0


@pytest.mark.parametrize("idx", [4, 5])
def test_safety_regression_every_cell(idx):
    """test/footprints.cpp:205-242: a lethal cell anywhere in the footprint, inflated by the split radius,
    must make the split check fail; the free map must pass."""
    res, radius = 0.05, 0.25
    cells = footprint_cells(FOOTPRINTS[idx], 0.5236, res)
    bank = [split.split_cells(cells, res, radius)]
    rad = int(math.ceil(radius / res))
    yy, xx = np.mgrid[-rad:rad + 1, -rad:rad + 1]
    disk = np.hypot(xx, yy) * res <= radius
    off = np.array([[100, 100]], dtype=np.int32)
    assert po.discrete_footprint_batch(po.Costmap(np.zeros((200, 200), np.uint8), res), bank, off).is_free[0] == 1
    misses = 0
    for cx, cy in cells[::5]:
        data = np.zeros((200, 200), dtype=np.uint8)
        x, y = cx + 100, cy + 100
        data[y - rad:y + rad + 1, x - rad:x + rad + 1][disk] = 253
        data[y, x] = 254
        misses += int(po.discrete_footprint_batch(po.Costmap(data, res), bank, off).is_free[0])
    assert misses == 0
