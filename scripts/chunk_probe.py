import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.getcwd())
import cover_b200
from cover_b200 import workloads as wl
ctx = cover_b200.Context(0)
data = wl.inflated_costmap(4000, 4000, 0.05, seed=3)
cmap = cover_b200.Costmap(ctx, data, 0.05)
lat = cover_b200.Lattice(**wl.dwa_lattice())
n = lat.count
pinned = torch.from_numpy(data).pin_memory().numpy()
bits = torch.empty((n + 31) // 32, dtype=torch.int32).pin_memory()
not_ok = torch.empty(1, dtype=torch.int64).pin_memory()
host_out = cover_b200.Result(free_bits=bits.numpy().view(np.uint32), not_ok=not_ok.numpy())
def both():
    cmap.upload_async(pinned)
    cmap.area_is_free(wl.RECT, lat, out=host_out)
for _ in range(3): both()
t0 = time.perf_counter()
for _ in range(10): both()
print(os.environ.get("COVER_B200_CHUNKS"), (time.perf_counter() - t0) / 10 * 1e3, "ms")
