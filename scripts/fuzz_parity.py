#!/usr/bin/env python
"""Large one-off parity fuzz (not part of the pytest suite): CUDA vs the CPU oracle on millions of poses.
usage: python scripts/fuzz_parity.py [poses_per_case]"""
import math, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cover_b200 as cb
from cover_b200 import workloads as wl
from oracle import pyoracle as po
from tests.fixtures import FOOTPRINTS, FOOTPRINT_NAMES, RECT

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
threads = po.hardware_threads()
ctx = cb.Context(0)
rng = np.random.default_rng(int(os.environ.get("FUZZ_SEED", "12345")))
data = wl.inflated_costmap(1500, 1300, 0.05, seed=5, obstacle_p=4e-4)
data[rng.random(data.shape) < 2e-4] = 255
cmap = cb.Costmap(ctx, data, 0.05, -1.0, 0.5)
omap = po.Costmap(data, 0.05, -1.0, 0.5)
ops = [("is_free", po.OP_IS_FREE), ("accumulate_cost", po.OP_ACCUMULATE), ("max_cost", po.OP_MAX)]
bad = 0
t0 = time.time()
shapes = [("rect", RECT)] + list(zip(FOOTPRINT_NAMES, FOOTPRINTS)) + [("potato48", wl.potato(48))]
for name, poly in shapes:
    m = n if poly.shape[1] <= 6 or name in ("TORU", "SOTO", "potato48") else n // 4
    poses = wl.random_poses(m, seed=hash(name) % 1000, lo=-2.0, hi=76.0)
    for shape, prefix in ((po.SHAPE_AREA, "area"), (po.SHAPE_OUTLINE, "outline")):
        for oname, op in ops:
            g = getattr(cmap, f"{prefix}_{oname}")(poly, poses)
            c = po.polygon_batch(omap, poly, poses, shape=shape, op=op, threads=threads)
            ok = np.array_equal(g.is_free, c.is_free) and np.array_equal(g.status, c.status) and (op == po.OP_IS_FREE or np.array_equal(g.cost, c.cost, equal_nan=True))
            bad += not ok
            print(f"{name:9s} {prefix:7s} {oname:15s} {m:8d} poses  free={c.is_free.mean():.3f}  {'OK' if ok else 'MISMATCH'}", flush=True)
# lattices at several pitches / headings
for pitch in (0.05, 0.0313, 0.11):
    th = rng.uniform(-math.pi, math.pi, 11)
    desc = dict(x0=-1.5, dx=pitch, nx=int(78 / pitch), y0=0.0, dy=pitch, ny=int(40 / pitch) // 4, cs=np.stack([np.cos(th), np.sin(th)], 1))
    for oname, op in ops:
        g = getattr(cmap, f"area_{oname}")(RECT, cb.Lattice(**desc))
        c = po.polygon_batch(omap, RECT, po.Lattice(**desc), op=op, threads=threads)
        ok = np.array_equal(g.is_free, c.is_free) and np.array_equal(g.status, c.status) and (op == po.OP_IS_FREE or np.array_equal(g.cost, c.cost, equal_nan=True))
        bad += not ok
        print(f"lattice pitch {pitch:6.4f} {oname:15s} {len(c.is_free):9d} poses free={c.is_free.mean():.3f} {'OK' if ok else 'MISMATCH'}", flush=True)
# random lattices x random polygons: the shape cache's span tables (convex / two pairs per row / split wide pairs / too
# complex -> general kernel), column groups, heading runs, sub-ranges that start and end inside a row
n_rand = int(os.environ.get("FUZZ_LATTICES", "40"))
for case in range(n_rand):
    nv = int(rng.integers(1, 13))
    radius = float(rng.choice([0.15, 0.4, 0.8, 1.6]))
    if rng.random() < 0.5:  # star-shaped around the origin (often non-convex), else an arbitrary vertex soup
        ang = np.sort(rng.uniform(0, 2 * math.pi, nv))
        rad = radius * rng.uniform(0.3, 1.0, nv)
        poly = np.stack([rad * np.cos(ang), rad * np.sin(ang)])
    else:
        poly = rng.uniform(-radius, radius, (2, nv))
    n_th = int(rng.integers(1, 7))
    th = np.where(rng.random(n_th) < 0.3, rng.integers(-2, 3, n_th) * (math.pi / 2), rng.uniform(-math.pi, math.pi, n_th))
    scale = 1.0 if rng.random() < 0.8 else float(rng.uniform(0.5, 1.5))  # |(c, s)| need not be 1
    pitch_x = float(rng.choice([0.05, 0.05, 0.0313, 0.1, 0.017])) * (1 if rng.random() < 0.8 else -1)
    pitch_y = float(rng.choice([0.05, 0.05, 0.0313, 0.1, 0.017])) * (1 if rng.random() < 0.8 else -1)
    nx, ny = int(rng.integers(1, 260)), int(rng.integers(1, 90))
    desc = dict(x0=float(rng.uniform(-2.0, 76.0)), dx=pitch_x, nx=nx, y0=float(rng.uniform(-2.0, 40.0)), dy=pitch_y, ny=ny,
                cs=scale * np.stack([np.cos(th), np.sin(th)], 1))
    total = nx * ny * n_th
    first = int(rng.integers(0, total)) if rng.random() < 0.4 else 0
    count = int(rng.integers(1, total - first + 1)) if rng.random() < 0.4 else total - first
    for shape, prefix in ((po.SHAPE_AREA, "area"), (po.SHAPE_OUTLINE, "outline")):
        for oname, op in ops:
            g = getattr(cmap, f"{prefix}_{oname}")(poly, cb.Lattice(**desc), first=first, count=count)
            c = po.polygon_batch(omap, poly, po.Lattice(**desc), shape=shape, op=op, threads=threads, first=first, count=count)
            ok = np.array_equal(g.is_free, c.is_free) and np.array_equal(g.status, c.status) and (op == po.OP_IS_FREE or np.array_equal(g.cost, c.cost, equal_nan=True))
            bad += not ok
            if not ok or oname == "is_free":
                print(f"random lattice {case:3d} nv={nv:2d} r={radius:4.2f} {nx:3d}x{ny:2d}x{n_th} [{first}, +{count}) {prefix:7s} {oname:15s} "
                      f"free={c.is_free.mean():.3f} ok_status={np.mean(c.status == 0):.2f} {'OK' if ok else 'MISMATCH'}", flush=True)
print(f"done in {time.time() - t0:.0f} s, mismatching cases: {bad}")
sys.exit(1 if bad else 0)
