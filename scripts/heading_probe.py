#!/usr/bin/env python
"""Per-heading cost of the C4 sweep: axis-aligned headings put the vertices on cell boundaries, where floating-point
noise in the lattice translation makes the discretised shape vary from pose to pose."""
import json, math, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cover_b200
from cover_b200 import workloads as wl

ctx = cover_b200.Context(0)
data = wl.inflated_costmap(4000, 4000, 0.05, seed=3)
cmap = cover_b200.Costmap(ctx, data, 0.05)
stream = torch.cuda.ExternalStream(ctx.stream)
window = int(os.environ.get("COVER_BENCH_WINDOW", "3"))
ALL = os.environ.get("ALL_HEADINGS") == "1"
for deg in ([-180.0 + 5.0 * k for k in range(72)] if ALL else (0.0, 5.0, 45.0, 90.0, 180.0, 37.0)):
    th = -math.pi + 2 * math.pi * round((deg + 180.0) / 5.0) / 72 if ALL else math.radians(deg)
    desc = wl.dwa_lattice(x0=20.0 + 2.0 * window, y0=20.0 + 3.0 * window, n_theta=1)
    desc["cs"] = np.array([[math.cos(th), math.sin(th)]] * 4)
    lat = cover_b200.Lattice(**desc)
    n = lat.count
    out = cover_b200.Result(torch.empty(n, dtype=torch.uint8, device="cuda"), None, torch.empty(n, dtype=torch.uint8, device="cuda"))
    for _ in range(2):
        cmap.area_is_free(wl.RECT, lat, out=out)
    ctx.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(stream)
    for _ in range(5):
        cmap.area_is_free(wl.RECT, lat, out=out)
    e.record(stream)
    ctx.synchronize()
    ms = s.elapsed_time(e) / 5
    print(json.dumps({"window": window, "heading_deg": deg, "poses": n, "ms": ms, "checks_per_s": n / ms * 1e3}), flush=True)
