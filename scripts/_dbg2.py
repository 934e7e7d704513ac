import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["COVER_B200_DEBUG_SYNC"]="1"
import numpy as np, cover_b200 as cb
from oracle import pyoracle as po
rng = np.random.default_rng(34)
data = rng.integers(0, 253, (256, 256), dtype=np.uint8)
omap = po.Costmap(data, 0.05)
th = -math.pi + 2 * math.pi * np.arange(12) / 12
cs = np.stack([np.cos(th), np.sin(th)], 1)
poly = np.array([[0.01], [0.01]])
import subprocess
if len(sys.argv) > 1:
    sel = [int(a) for a in sys.argv[1].split(",")]
    nx, ny = int(sys.argv[2]), int(sys.argv[3])
    desc = dict(x0=0.2, dx=0.05, nx=nx, y0=0.3, dy=0.05, ny=ny, cs=cs[sel])
    ctx = cb.Context(0)
    cmap = cb.Costmap(ctx, data, 0.05)
    try:
        gpu = cmap.area_is_free(poly, cb.Lattice(**desc))
        cpu = po.polygon_batch(omap, poly, po.Lattice(**desc), threads=8)
        print(sel, nx, ny, "ok, mismatches", int((gpu.is_free != cpu.is_free).sum()), int((gpu.status != cpu.status).sum()))
    except Exception as e:
        print(sel, nx, ny, "ERR", str(e)[:80])
else:
    for sel in ["0", "3", "6", "9", "0,1", "0,1,2,3,4,5", "6,7,8,9,10,11", "0,1,2,3,4,5,6,7,8,9,10,11"]:
        for nx, ny in ((230, 100), (128, 100), (230, 64)):
            subprocess.run([sys.executable, __file__, sel, str(nx), str(ny)])
