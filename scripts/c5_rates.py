"""C5 shapes (bench.py's c5_shapes) x random poses on the 8000^2 inflated map: rate per shape and per fold, parity of
the first 50 000 poses against the CPU side.  Usage: python scripts/c5_rates.py [n_poses] [op ...]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import cover_b200  # noqa: E402
from cover_b200 import workloads as wl  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
ops = sys.argv[2:] or ["is_free"]
ctx = cover_b200.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream)
data = wl.inflated_costmap(8000, 8000, 0.05, seed=7)
cmap = cover_b200.Costmap(ctx, data, 0.05)
omap = po.Costmap(data, 0.05)
poses_np = wl.random_poses(n, seed=8, lo=8.0, hi=392.0)
poses = torch.from_numpy(poses_np).cuda()
backend = po.reference() or po.default()
m = 50_000
for op in ops:
    value = op != "is_free"
    out = cover_b200.Result(torch.empty(n, dtype=torch.uint8, device="cuda"), torch.empty(n, dtype=torch.float64, device="cuda") if value else None,
                            torch.empty(n, dtype=torch.uint8, device="cuda"))
    code = {"is_free": po.OP_IS_FREE, "accumulate_cost": po.OP_ACCUMULATE, "max_cost": po.OP_MAX}[op]
    for name, poly in bench.c5_shapes():
        fn = getattr(cmap, f"area_{op}")
        n0 = ctx.launch_count
        ms = bench.device_timed(ctx, stream, lambda: fn(poly, poses, out=out), 3, warm=1)
        launches = (ctx.launch_count - n0) // 4
        r = backend.polygon_batch(omap, poly, poses_np[:m], op=code, threads=16)
        if value:
            ok = (r.status == 0) & (out.status[:m].cpu().numpy() == 0)
            bad = int((out.cost[:m].cpu().numpy()[ok] != r.cost[ok]).sum()) + int(((out.status[:m].cpu().numpy() == 1) != (r.status == 1)).sum())
        else:
            bad = bench.compare(out.is_free[:m].cpu().numpy(), out.status[:m].cpu().numpy(), r.is_free, r.status)
        print(f"{op:16s} {name:12s} {ms:8.3f} ms  {n / ms / 1e6:8.2f} G/s  launches {launches}  mismatches {bad}  not_ok {int((out.status != 0).sum())}", flush=True)
