"""tests/golden/reference_outputs.npz holds outputs of the REFERENCE ITSELF (magazino/cover's own sources
compiled over shim headers, see tests/golden/make_golden.py).  The CPU restatement must reproduce them here;
the CUDA kernels must reproduce them on the B200 (`-m gpu`)."""
import os

import numpy as np
import pytest

from oracle import pyoracle as po
from tests.fixtures import FOOTPRINTS, RECT
from tests.golden.make_golden import inputs

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_outputs.npz"))
INP = inputs()
POLYS = [RECT] + [f * 0.35 for f in FOOTPRINTS]
TORU = FOOTPRINTS[5]
OUTLINES, AREAS = [TORU * 0.4], [TORU, RECT * 0.5 + np.array([[0.2], [0.1]])]


def same_cost(a, b):
    np.testing.assert_array_equal(np.nan_to_num(a, nan=-99.0), np.nan_to_num(b, nan=-99.0))


def test_fixture_is_not_trivial():
    assert 0.05 < GOLD["poly0_area_free"].mean() < 0.95
    assert GOLD["poly1_area_logic_error"].sum() == 0
    assert len(GOLD["s_area_cells"]) > 1000


def test_port_reproduces_the_reference_outputs():
    cmap = po.Costmap(INP["data"], INP["res"], INP["ox"], INP["oy"])
    port = po.default()
    for k, poly in enumerate(POLYS):
        for shape, sname in ((po.SHAPE_AREA, "area"), (po.SHAPE_OUTLINE, "outline")):
            r = port.polygon_batch(cmap, poly, INP["poses"], shape=shape)
            a = port.polygon_batch(cmap, poly, INP["poses"], shape=shape, op=po.OP_ACCUMULATE)
            np.testing.assert_array_equal(r.is_free, GOLD[f"poly{k}_{sname}_free"])
            np.testing.assert_array_equal((a.status == po.LOGIC_ERROR).astype(np.uint8), GOLD[f"poly{k}_{sname}_logic_error"])
            same_cost(a.cost, GOLD[f"poly{k}_{sname}_cost"])
    np.testing.assert_array_equal(port.rays_batch(cmap, INP["segs"]).is_free, GOLD["ray_free"])
    np.testing.assert_array_equal(port.rays_batch(cmap, INP["segs"], offset=(5, -3), op=po.OP_ACCUMULATE).cost, GOLD["ray_cost_offset"])
    np.testing.assert_array_equal(port.continuous_footprint_batch(cmap, OUTLINES, AREAS, INP["poses"]).is_free, GOLD["split_free"])
    oc, ac = port.make_discrete_footprint(INP["res"], OUTLINES, AREAS)
    np.testing.assert_array_equal(oc, GOLD["discrete_outline_cells"])
    np.testing.assert_array_equal(ac, GOLD["discrete_area_cells"])
    np.testing.assert_array_equal(port.discrete_footprint_batch(cmap, [(oc, ac)], INP["offs"]).is_free, GOLD["discrete_free"])
    np.testing.assert_array_equal(port.cells_batch(cmap, ac, INP["offs"]).is_free, GOLD["cells_free"])
    sparse = port.to_discrete(0.05, FOOTPRINTS[0])
    np.testing.assert_array_equal(port.outline_cells(sparse), GOLD["s_outline_cells"])
    np.testing.assert_array_equal(port.area_cells([GOLD["s_outline_cells"]]), GOLD["s_area_cells"])


@pytest.mark.gpu
def test_cuda_reproduces_the_reference_outputs():
    import cover_b200 as cb
    ctx = cb.Context(0)
    try:
        cmap = cb.Costmap(ctx, INP["data"], INP["res"], INP["ox"], INP["oy"])
        for k, poly in enumerate(POLYS):
            for sname in ("area", "outline"):
                r = getattr(cmap, f"{sname}_is_free")(poly, INP["poses"])
                a = getattr(cmap, f"{sname}_accumulate_cost")(poly, INP["poses"])
                np.testing.assert_array_equal(r.is_free, GOLD[f"poly{k}_{sname}_free"])
                np.testing.assert_array_equal((a.status == cb.STATUS_LOGIC_ERROR).astype(np.uint8), GOLD[f"poly{k}_{sname}_logic_error"])
                same_cost(a.cost, GOLD[f"poly{k}_{sname}_cost"])
        np.testing.assert_array_equal(cmap.ray_is_free(INP["segs"]).is_free, GOLD["ray_free"])
        np.testing.assert_array_equal(cmap.ray_accumulate_cost(INP["segs"], offset=(5, -3)).cost, GOLD["ray_cost_offset"])
        np.testing.assert_array_equal(cb.SplitFootprint(ctx, OUTLINES, AREAS).is_free(cmap, INP["poses"]).is_free, GOLD["split_free"])
        bank = cb.FootprintBank(ctx, [(GOLD["discrete_outline_cells"], GOLD["discrete_area_cells"])])
        np.testing.assert_array_equal(bank.is_free(cmap, INP["offs"]).is_free, GOLD["discrete_free"])
        np.testing.assert_array_equal(cb.DiscretePolygon(ctx, GOLD["discrete_area_cells"]).is_free(cmap, INP["offs"]).is_free, GOLD["cells_free"])
        # expand: to_outline / to_area of the S-shape, cell for cell (pose = identity, origin 0)
        flat = cb.Costmap(ctx, np.zeros((8, 8), dtype=np.uint8), 0.05)
        ident = np.array([[1.0, 0.0, 0.0, 0.0]])
        np.testing.assert_array_equal(flat.outline_cells(FOOTPRINTS[0], ident)[0], GOLD["s_outline_cells"])
        np.testing.assert_array_equal(flat.area_cells(FOOTPRINTS[0], ident)[0], GOLD["s_area_cells"])
    finally:
        ctx.close()
