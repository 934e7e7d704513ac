import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["COVER_B200_DEBUG_SYNC"]="1"
import numpy as np, cover_b200 as cb, torch
from oracle import pyoracle as po
from tests.fixtures import RECT, FOOTPRINTS
from cover_b200 import workloads as wl
def mixed_map(rng, sx, sy, lethal=0.002):
    data = rng.integers(0, 253, (sy, sx), dtype=np.uint8)
    r = rng.random((sy, sx))
    data[r < lethal] = 254
    data[(r >= lethal) & (r < 1.5 * lethal)] = 255
    data[(r >= 1.5 * lethal) & (r < 2.5 * lethal)] = 253
    return data
ctx = cb.Context(0)
rng = np.random.default_rng(34)
data = mixed_map(rng, 256, 256, lethal=0.004)
cmap = cb.Costmap(ctx, data, 0.05)
omap = po.Costmap(data, 0.05)
th = -math.pi + 2 * math.pi * np.arange(12) / 12
desc = dict(x0=0.2, dx=0.05, nx=230, y0=0.3, dy=0.05, ny=100, cs=np.stack([np.cos(th), np.sin(th)], 1))
polys = [np.array([[0.4, -0.3, -0.3], [0.0, 0.25, -0.25]]), FOOTPRINTS[6] * 0.8, np.array([[0.01], [0.01]]),
         np.array([[0.0, 0.3], [0.0, 0.45]]), np.array([[0.6, 0.6, -0.6, -0.6], [0.05, -0.05, -0.05, 0.05]])]
for pi, poly in enumerate(polys):
    try:
        gpu = cmap.area_is_free(poly, cb.Lattice(**desc))
    except Exception as e:
        print("poly", pi, "ERR", e); break
    cpu = po.polygon_batch(omap, poly, po.Lattice(**desc), threads=8)
    bad = np.nonzero(gpu.is_free != cpu.is_free)[0]
    print("poly", pi, "mismatches", bad.size, "status mism", int((gpu.status != cpu.status).sum()))
    if bad.size:
        it = bad // (230*100); r = bad % (230*100); iy = r // 230; ix = r % 230
        print("  headings", np.unique(it), "iy range", iy.min(), iy.max(), "ix range", ix.min(), ix.max(), "gpu says free", int(gpu.is_free[bad].sum()))
        print("  first", list(zip(it[:10].tolist(), iy[:10].tolist(), ix[:10].tolist())))
# C4 bytes
data = wl.inflated_costmap(4000, 4000, 0.05, seed=3)
d4 = wl.dwa_lattice()
n = d4["nx"] * d4["ny"] * d4["cs"].shape[0]
for strip in ("2", "4", "8"):
    os.environ["COVER_B200_WORDS_STRIP"] = strip
    c2 = cb.Context(0)
    cm = cb.Costmap(c2, data, 0.05)
    out = cb.Result(torch.empty(n, dtype=torch.uint8, device="cuda"), None, torch.empty(n, dtype=torch.uint8, device="cuda"))
    try:
        cm.area_is_free(wl.RECT, cb.Lattice(**d4), out=out); c2.synchronize(); print("C4 bytes strip", strip, "ok", float(out.is_free.float().mean()))
    except Exception as e:
        print("C4 bytes strip", strip, "ERR", e); break
