"""One C4 lattice sweep configuration, a few steps, for ncu (launch list / --set full captures).
Usage: python scripts/prof_lattice.py [bits|bytes] [steps] [shape]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cover_b200  # noqa: E402
from cover_b200 import workloads as wl  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "bits"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
data = wl.inflated_costmap(4000, 4000, 0.05, seed=3)
desc = wl.dwa_lattice()
n = desc["nx"] * desc["ny"] * desc["cs"].shape[0]
map_dev = torch.from_numpy(data).cuda()
ctx = cover_b200.Context(0)
cmap = cover_b200.Costmap(ctx, data, 0.05)
lat = cover_b200.Lattice(**desc)
if mode == "bytes":
    out = cover_b200.Result(torch.empty(n, dtype=torch.uint8, device="cuda"), None, torch.empty(n, dtype=torch.uint8, device="cuda"))
else:
    out = cover_b200.Result(free_bits=torch.zeros((n + 31) // 32, dtype=torch.int32, device="cuda"), not_ok=torch.zeros(1, dtype=torch.int64, device="cuda"))
for _ in range(steps):
    cmap.upload(map_dev)
    cmap.area_is_free(wl.RECT, lat, out=out)
ctx.synchronize()
print("done", ctx.launch_count)
