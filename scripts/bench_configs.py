#!/usr/bin/env python
"""Secondary measurements: the other BASELINE.json configs (C1, C2, C3, C5) on one GPU, device-resident
inputs, CUDA events on the library's stream.  Not the driver's bench (that is bench.py = C4); the numbers
land in profiles/ as context for DESIGN.md.  Sizes can be scaled with --scale (1.0 = BASELINE sizes)."""
import argparse
import json
import math
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cover_b200  # noqa: E402
from cover_b200 import workloads as wl  # noqa: E402
from tests.fixtures import FOOTPRINTS  # noqa: E402


def timed(ctx, fn, reps=5, warm=2):
    stream = torch.cuda.ExternalStream(ctx.stream)
    for _ in range(warm):
        fn()
    ctx.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(stream)
    for _ in range(reps):
        fn()
    e.record(stream)
    ctx.synchronize()
    return s.elapsed_time(e) / reps


def dev_result(n, cost=False):
    return cover_b200.Result(torch.empty(n, dtype=torch.uint8, device="cuda"),
                             torch.empty(n, dtype=torch.float64, device="cuda") if cost else None,
                             torch.empty(n, dtype=torch.uint8, device="cuda"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    ctx = cover_b200.Context(0)
    rows = []

    def report(name, n, ms, extra=None):
        row = {"config": name, "poses": n, "ms": ms, "checks_per_s": n / (ms * 1e-3)}
        row.update(extra or {})
        rows.append(row)
        print(json.dumps(row), flush=True)

    want = lambda k: not args.only or k in args.only.split(",")

    if want("C1"):
        n = int(1e5 * max(args.scale, 1.0) * 10)  # 1e6 poses so that the launch is not latency-bound
        data = wl.random_costmap(1000, 1000, seed=1, lethal_p=1e-3)
        cmap = cover_b200.Costmap(ctx, data, 0.05)
        poses = torch.from_numpy(wl.random_poses(n, seed=2, lo=1.0, hi=49.0)).cuda()
        out = dev_result(n, cost=True)
        report("C1 rect area is_free, random poses, 1000^2", n, timed(ctx, lambda: cmap.area_is_free(wl.RECT, poses, out=out)),
               {"free_fraction": float(out.is_free.float().mean())})
        report("C1 rect area accumulate_cost, random poses, 1000^2", n, timed(ctx, lambda: cmap.area_accumulate_cost(wl.RECT, poses, out=out)))

    if want("C2"):
        n = int(1e6 * args.scale)
        data = wl.inflated_costmap(2000, 2000, 0.05, seed=3)
        cmap = cover_b200.Costmap(ctx, data, 0.05)
        poses = torch.from_numpy(wl.random_poses(n, seed=4, lo=1.0, hi=99.0)).cuda()
        out = dev_result(n, cost=True)
        for name, poly in (("rect", wl.RECT), ("TORU", FOOTPRINTS[5])):
            report(f"C2 {name} outline is_free, 2000^2 inflated", n, timed(ctx, lambda: cmap.outline_is_free(poly, poses, out=out)))
            report(f"C2 {name} outline accumulate_cost, 2000^2 inflated", n, timed(ctx, lambda: cmap.outline_accumulate_cost(poly, poses, out=out)))
        rng = np.random.default_rng(5)
        b = rng.integers(30, 1970, (n, 2))
        segs = torch.from_numpy(np.concatenate([b, b + rng.integers(-25, 26, (n, 2))], 1).astype(np.int32)).cuda()
        report("C2 rays (<= 25 cells) is_free, 2000^2 inflated", n, timed(ctx, lambda: cmap.ray_is_free(segs, out=out)))

    if want("C3"):
        from oracle import pyoracle as po  # host precompute of the 72 discrete footprints (stand-in for the caller's cache)
        n = int(1e7 * args.scale)
        data = wl.inflated_costmap(4000, 4000, 0.05, seed=3)
        cmap = cover_b200.Costmap(ctx, data, 0.05)
        pot = wl.potato(48)
        outlines, areas = [wl.scaled(pot, 0.45)], [pot]   # a hand-made split: inner skeleton outline + full remainder scan
        bank = []
        for k in range(72):
            R = wl.rotation(-math.pi + 2 * math.pi * k / 72)
            bank.append(po.make_discrete_footprint(0.05, [R @ o for o in outlines], [R @ a for a in areas]))
        fb = cover_b200.FootprintBank(ctx, bank)
        rng = np.random.default_rng(5)
        offs = torch.from_numpy(rng.integers(40, 3960, (n, 2)).astype(np.int32)).cuda()
        which = torch.from_numpy(rng.integers(0, 72, n).astype(np.int32)).cuda()
        out = dev_result(n)
        report("C3 potato discrete_footprint bank (72 headings) is_free, 4000^2", n, timed(ctx, lambda: fb.is_free(cmap, offs, which, out=out), reps=3),
               {"cells_per_footprint": int(np.mean([len(o) + len(a) for o, a in bank])), "free_fraction": float(out.is_free.float().mean())})
        sf = cover_b200.SplitFootprint(ctx, outlines, areas)
        n2 = max(n // 10, 1000)
        poses = torch.from_numpy(wl.random_poses(n2, seed=6, lo=2.0, hi=198.0)).cuda()
        out2 = dev_result(n2)
        report("C3 potato continuous_footprint is_free (re-rasterised per pose), 4000^2", n2, timed(ctx, lambda: sf.is_free(cmap, poses, out=out2), reps=3))

    if want("C5"):
        n = int(1e7 * args.scale)
        data = wl.inflated_costmap(8000, 8000, 0.05, seed=7)
        cmap = cover_b200.Costmap(ctx, data, 0.05)
        poses = torch.from_numpy(wl.random_poses(n, seed=8, lo=8.0, hi=392.0)).cuda()
        out = dev_result(n)
        shapes = [("rect(4)", wl.RECT), ("L(6)", FOOTPRINTS[6]), ("S(6)", FOOTPRINTS[0]), ("W(8)", FOOTPRINTS[1]), ("M(10)", FOOTPRINTS[2]),
                  ("SOTO(22)", FOOTPRINTS[4]), ("TORU(32)", FOOTPRINTS[5]), ("potato(64)", wl.potato(64))]
        for name, poly in shapes:
            m = n if poly.shape[1] <= 4 or name.startswith(("L", "TORU", "SOTO", "potato")) else n // 10
            sub = poses[:m]
            o = cover_b200.Result(out.is_free[:m], None, out.status[:m])
            report(f"C5 {name} area is_free, random poses, 8000^2", m, timed(ctx, lambda: cmap.area_is_free(poly, sub, out=o), reps=2, warm=1),
                   {"free_fraction": float(o.is_free.float().mean()), "not_ok": int((o.status != 0).sum())})
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "bench_configs.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
