"""Device time of cover_map_inflate on the C4 map (4000^2, point obstacles w.p. 2e-4, radius 20 cells) and on a denser one."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cover_b200
from cover_b200 import workloads as wl
ctx = cover_b200.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream)
rad, lut = wl.inflation_lut(0.05)
for p in (2e-4, 5e-3):
    rng = np.random.default_rng(3)
    raw = np.zeros((4000, 4000), dtype=np.uint8)
    raw[rng.random((4000, 4000)) < p] = 254
    dev = torch.from_numpy(raw).cuda()
    cmap = cover_b200.Costmap(ctx, raw, 0.05)
    ts = []
    for _ in range(6):
        cmap.upload(dev)
        ctx.synchronize()
        t0 = time.perf_counter()
        cmap.inflate(rad, lut)   # synchronous call
        ts.append(time.perf_counter() - t0)
    print(f"obstacle density {p}: inflate 4000^2 radius {rad}: {min(ts) * 1e3:.3f} ms (host-timed, synchronous call incl. the LUT copy)")
