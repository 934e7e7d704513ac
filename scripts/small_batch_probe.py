#!/usr/bin/env python
"""Latency of small lattice sweeps (what one DWA cycle of a real planner looks like): device time per call and wall
time per call with host results, for growing (x, y, heading) lattices on the C4 map."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cover_b200
from cover_b200 import workloads as wl

ctx = cover_b200.Context(0)
data = wl.inflated_costmap(4000, 4000, 0.05, seed=3)
cmap = cover_b200.Costmap(ctx, data, 0.05)
stream = torch.cuda.ExternalStream(ctx.stream)
for nx, ny, nt in ((20, 20, 24), (40, 40, 72), (100, 100, 72), (300, 300, 72), (600, 600, 72)):
    lat = cover_b200.Lattice(**wl.dwa_lattice(nx=nx, ny=ny, n_theta=nt, x0=60.0, y0=70.0))
    n = lat.count
    dev = cover_b200.Result(torch.empty(n, dtype=torch.uint8, device="cuda"))
    for _ in range(3):
        cmap.area_is_free(wl.RECT, lat, out=dev)
    ctx.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(stream)
    for _ in range(20):
        cmap.area_is_free(wl.RECT, lat, out=dev)
    e.record(stream)
    ctx.synchronize()
    dev_us = s.elapsed_time(e) / 20 * 1e3
    t0 = time.perf_counter()
    for _ in range(20):
        host = cmap.area_is_free(wl.RECT, lat)
    wall_us = (time.perf_counter() - t0) / 20 * 1e6
    print(json.dumps({"lattice": [nx, ny, nt], "poses": n, "device_us": dev_us, "wall_us_host_results": wall_us, "checks_per_s_device": n / dev_us * 1e6}), flush=True)
