#!/usr/bin/env python
"""C4 lattice with the other folds (accumulate_cost, max_cost) and the outline generator: context numbers."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cover_b200
from cover_b200 import workloads as wl

ctx = cover_b200.Context(0)
data = wl.inflated_costmap(4000, 4000, 0.05, seed=3)
cmap = cover_b200.Costmap(ctx, data, 0.05)
desc = wl.dwa_lattice(n_theta=int(sys.argv[1]) if len(sys.argv) > 1 else 18)
lat = cover_b200.Lattice(**desc)
n = lat.count
out = cover_b200.Result(torch.empty(n, dtype=torch.uint8, device="cuda"), torch.empty(n, dtype=torch.float64, device="cuda"), torch.empty(n, dtype=torch.uint8, device="cuda"))
stream = torch.cuda.ExternalStream(ctx.stream)
sys.path.insert(0, os.path.join(ROOT, "tests"))
poly_name = os.environ.get("POLY", "rect")
if poly_name == "rect":
    POLY = wl.RECT
elif poly_name == "rect2":
    POLY = 2.0 * wl.RECT                   # 2.0 x 1.2 m: 40 x 24 cells, pairs wider than 32 cells are split
elif poly_name == "potato":
    POLY = wl.scaled(wl.potato(16), 0.6)   # 16 vertices, ~0.7 m across: hashed shape keys
else:
    from fixtures import FOOTPRINTS, FOOTPRINT_NAMES
    POLY = FOOTPRINTS[FOOTPRINT_NAMES.index(poly_name)]
for name in ("area_is_free", "area_accumulate_cost", "area_max_cost", "outline_is_free", "outline_accumulate_cost"):
    fn = getattr(cmap, name)
    for _ in range(2):
        fn(POLY, lat, out=out)
    ctx.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(stream)
    for _ in range(3):
        fn(POLY, lat, out=out)
    e.record(stream)
    ctx.synchronize()
    ms = s.elapsed_time(e) / 3
    print(json.dumps({"polygon": poly_name, "op": name, "poses": n, "ms": ms, "checks_per_s": n / ms * 1e3}), flush=True)
