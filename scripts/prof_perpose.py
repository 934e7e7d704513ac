"""Per-pose kernels, one launch each at BASELINE-like sizes, for ncu captures (profiles/ncu_r02_*):
  area_pair  C1 rect, random poses (k_area_pair)      area_rows  C5 S-shape, random poses (k_area_rows)
  outline    C2 rect outline (k_outline)               footprints C3 72-heading bank x offsets (k_footprints)
  split      C3 split footprint, pose form (k_split)
Usage: python scripts/prof_perpose.py <which> [n_poses]"""
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cover_b200  # noqa: E402
from cover_b200 import workloads as wl  # noqa: E402
from tests.fixtures import FOOTPRINTS  # noqa: E402

which = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
ctx = cover_b200.Context(0)
out = cover_b200.Result(torch.empty(n, dtype=torch.uint8, device="cuda"), None, torch.empty(n, dtype=torch.uint8, device="cuda"))
stream = torch.cuda.ExternalStream(ctx.stream)


def run(fn, reps=3):
    for _ in range(2):
        fn()
    ctx.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(stream)
    for _ in range(reps):
        fn()
    e.record(stream)
    ctx.synchronize()
    ms = s.elapsed_time(e) / reps
    print(f"{which}: {n} poses, {ms:.4f} ms, {n / ms / 1e6:.1f} M checks/s")


if which.startswith("area_"):
    big = which != "area_pair"
    size = 8000 if big else 1000
    data = wl.inflated_costmap(size, size, 0.05, seed=7) if big else wl.random_costmap(1000, 1000, seed=1, lethal_p=1e-3)
    cmap = cover_b200.Costmap(ctx, data, 0.05)
    poses = torch.from_numpy(wl.random_poses(n, seed=8, lo=8.0, hi=size * 0.05 - 8.0)).cuda()
    poly = {"area_pair": wl.RECT, "area_rows": FOOTPRINTS[0], "area_M": FOOTPRINTS[2], "area_W": FOOTPRINTS[1], "area_SOTO": FOOTPRINTS[4],
            "area_L": FOOTPRINTS[6], "area_TORU": FOOTPRINTS[5], "area_potato": wl.potato(64)}[which]
    run(lambda: cmap.area_is_free(poly, poses, out=out))
elif which == "outline":
    data = wl.inflated_costmap(2000, 2000, 0.05, seed=3)
    cmap = cover_b200.Costmap(ctx, data, 0.05)
    poses = torch.from_numpy(wl.random_poses(n, seed=4, lo=1.0, hi=99.0)).cuda()
    run(lambda: cmap.outline_is_free(wl.RECT, poses, out=out))
else:
    data = wl.inflated_costmap(4000, 4000, 0.05, seed=3)
    cmap = cover_b200.Costmap(ctx, data, 0.05)
    pot = wl.potato(48)
    outlines, areas = [wl.scaled(pot, 0.45)], [pot]
    if which == "footprints":
        th = -math.pi + 2 * math.pi * np.arange(72) / 72
        head = np.stack([np.cos(th), np.sin(th), np.zeros(72), np.zeros(72)], 1)
        zero_map = cover_b200.Costmap(ctx, np.zeros((8, 8), dtype=np.uint8), 0.05)
        oc, os_, _ = zero_map.outline_cells(outlines[0], head)
        ac, as_, _ = zero_map.area_cells(areas[0], head)
        fb = cover_b200.FootprintBank(ctx, [(oc[os_[k]:os_[k + 1]], ac[as_[k]:as_[k + 1]]) for k in range(72)])
        rng = np.random.default_rng(5)
        offs = torch.from_numpy(rng.integers(40, 3960, (n, 2)).astype(np.int32)).cuda()
        sel = torch.from_numpy(rng.integers(0, 72, n).astype(np.int32)).cuda()
        run(lambda: fb.is_free(cmap, offs, sel, out=out))
    else:
        sf = cover_b200.SplitFootprint(ctx, outlines, areas)
        poses = torch.from_numpy(wl.random_poses(n, seed=6, lo=2.0, hi=198.0)).cuda()
        run(lambda: sf.is_free(cmap, poses, out=out))
