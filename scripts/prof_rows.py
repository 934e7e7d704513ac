import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
import cover_b200
from cover_b200 import workloads as wl
from tests.fixtures import FOOTPRINTS
ctx = cover_b200.Context(0)
data = wl.inflated_costmap(4000, 4000, 0.05, seed=7)
cmap = cover_b200.Costmap(ctx, data, 0.05)
poses = torch.from_numpy(wl.random_poses(300000, seed=8, lo=8.0, hi=192.0)).cuda()
out = cover_b200.Result(torch.empty(300000, dtype=torch.uint8, device="cuda"), None, torch.empty(300000, dtype=torch.uint8, device="cuda"))
for _ in range(3):
    cmap.area_is_free(FOOTPRINTS[0], poses, out=out)
ctx.synchronize()
print(float(out.is_free.float().mean()))
