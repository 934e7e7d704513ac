"""COVER_B200_TIMELINE=1 python scripts/timeline_probe.py: where the time of one headline step goes (events between the
library's operations on its main stream, printed by cover_ctx_synchronize)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import cover_b200  # noqa: E402
from cover_b200 import workloads as wl  # noqa: E402

data = bench.make_map()
desc = bench.lattice_for_rank(0)
n = desc["nx"] * desc["ny"] * desc["cs"].shape[0]
ctx = cover_b200.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream)
lat = cover_b200.Lattice(**desc)
cmap = cover_b200.Costmap(ctx, data, bench.RESOLUTION)
out = cover_b200.Result(free_bits=torch.zeros((n + 31) // 32, dtype=torch.int32, device="cuda"), not_ok=torch.zeros(1, dtype=torch.int64, device="cuda"))
map_dev = torch.from_numpy(data).cuda()
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
for k in range(6):
    with torch.cuda.stream(stream):
        flush.fill_(k)
        cmap.upload(map_dev)
        cmap.area_is_free(wl.RECT, lat, out=out)
    if k >= 3:
        print(f"--- step {k}", file=sys.stderr, flush=True)
    else:
        os.environ["QUIET"] = "1"
    ctx.synchronize()
