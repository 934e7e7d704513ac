#!/usr/bin/env python
"""Generates tests/golden/reference_outputs.npz from the REFERENCE ITSELF: magazino/cover's own base.cpp /
generators.cpp / costmap.cpp / footprints.cpp compiled unmodified over the shim headers into
oracle/_ref/libcover_ref.so (oracle/Makefile; needs /root/reference).  The fixture travels with the repo, so
the restatement (CPU tests) and the CUDA kernels (GPU tests) are checked against real reference outputs even
where oracle/_ref cannot be rebuilt.

    python tests/golden/make_golden.py        # rewrites reference_outputs.npz (deterministic: seeded inputs)
"""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402
from tests.fixtures import FOOTPRINTS, RECT  # noqa: E402

SEED = 20251019


def inputs():
    """Seeded inputs shared by the generator and by the tests that consume the fixture."""
    rng = np.random.default_rng(SEED)
    data = rng.integers(0, 253, (220, 260), dtype=np.uint8)
    r = rng.random(data.shape)
    data[r < 0.002] = 254
    data[(r >= 0.002) & (r < 0.003)] = 255
    data[(r >= 0.003) & (r < 0.005)] = 253
    th = rng.uniform(-math.pi, math.pi, 400)
    poses = np.stack([np.cos(th), np.sin(th), rng.uniform(-0.5, 11.5, 400), rng.uniform(-0.5, 9.5, 400)], axis=1)
    segs = rng.integers(-8, 270, (400, 4)).astype(np.int32)
    offs = rng.integers(-15, 250, (400, 2)).astype(np.int32)
    return dict(data=data, res=0.05, ox=-0.75, oy=0.4, poses=poses, segs=segs, offs=offs)


def main():
    ref = po.reference()
    if ref is None:
        raise SystemExit("oracle/_ref/libcover_ref.so is missing: run `make -C oracle ref` where /root/reference exists")
    inp = inputs()
    cmap = po.Costmap(inp["data"], inp["res"], inp["ox"], inp["oy"])
    out = {}
    polys = [RECT] + [f * 0.35 for f in FOOTPRINTS]
    for k, poly in enumerate(polys):
        for shape, sname in ((po.SHAPE_AREA, "area"), (po.SHAPE_OUTLINE, "outline")):
            r = ref.polygon_batch(cmap, poly, inp["poses"], shape=shape, op=po.OP_IS_FREE)
            a = ref.polygon_batch(cmap, poly, inp["poses"], shape=shape, op=po.OP_ACCUMULATE)
            out[f"poly{k}_{sname}_free"] = r.is_free
            out[f"poly{k}_{sname}_cost"] = a.cost
            out[f"poly{k}_{sname}_logic_error"] = (a.status == po.LOGIC_ERROR).astype(np.uint8)
    out["ray_free"] = ref.rays_batch(cmap, inp["segs"]).is_free
    out["ray_cost_offset"] = ref.rays_batch(cmap, inp["segs"], offset=(5, -3), op=po.OP_ACCUMULATE).cost
    toru = FOOTPRINTS[5]
    outlines, areas = [toru * 0.4], [toru, RECT * 0.5 + np.array([[0.2], [0.1]])]
    out["split_free"] = ref.continuous_footprint_batch(cmap, outlines, areas, inp["poses"]).is_free
    oc, ac = ref.make_discrete_footprint(inp["res"], outlines, areas)
    out["discrete_outline_cells"], out["discrete_area_cells"] = oc, ac
    out["discrete_free"] = ref.discrete_footprint_batch(cmap, [(oc, ac)], inp["offs"]).is_free
    out["cells_free"] = ref.cells_batch(cmap, ac, inp["offs"]).is_free
    # dense generators: the S-shape outline and area at 0.05 m (order matters)
    sparse = ref.to_discrete(0.05, FOOTPRINTS[0])
    out["s_outline_cells"] = ref.outline_cells(sparse)
    out["s_area_cells"] = ref.area_cells([out["s_outline_cells"]])
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_outputs.npz"), **out)
    print("wrote reference_outputs.npz:", {k: v.shape for k, v in out.items() if k.endswith("free")})


if __name__ == "__main__":
    main()
