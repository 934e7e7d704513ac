"""Device-timed C4 lattice sweep (map resident, planes rebuilt per step) for the strip widths of k_lw_sweep and the
shape-cache kernel, byte and bit-packed outputs.  Usage: python scripts/lattice_probe.py [reps]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cover_b200  # noqa: E402
from cover_b200 import workloads as wl  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
data = wl.inflated_costmap(4000, 4000, 0.05, seed=3)
desc = wl.dwa_lattice()
n = desc["nx"] * desc["ny"] * desc["cs"].shape[0]
map_dev = torch.from_numpy(data).cuda()
ref = None
for env in ({"COVER_B200_NO_WORDS": "1"}, {"COVER_B200_WORDS_STRIP": "2"}, {"COVER_B200_WORDS_STRIP": "4"}, {"COVER_B200_WORDS_STRIP": "8"}):
    for k in ("COVER_B200_NO_WORDS", "COVER_B200_WORDS_STRIP"):
        os.environ.pop(k, None)
    os.environ.update(env)
    ctx = cover_b200.Context(0)
    stream = torch.cuda.ExternalStream(ctx.stream)
    cmap = cover_b200.Costmap(ctx, data, 0.05)
    lat = cover_b200.Lattice(**desc)
    for mode in ("bytes", "bits"):
        if mode == "bytes":
            out = cover_b200.Result(torch.empty(n, dtype=torch.uint8, device="cuda"), None, torch.empty(n, dtype=torch.uint8, device="cuda"))
        else:
            out = cover_b200.Result(free_bits=torch.zeros((n + 31) // 32, dtype=torch.int32, device="cuda"), not_ok=torch.zeros(1, dtype=torch.int64, device="cuda"))
        for keep_planes in (False, True):
            times = []
            for r in range(reps + 3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                with torch.cuda.stream(stream):
                    e0.record(stream)
                    if not keep_planes:
                        cmap.upload(map_dev)
                    cmap.area_is_free(wl.RECT, lat, out=out)
                    e1.record(stream)
                ctx.synchronize()
                if r >= 3:
                    times.append(e0.elapsed_time(e1))
            if mode == "bytes":
                got = out.is_free.cpu().numpy()
            else:
                got = np.unpackbits(out.free_bits.cpu().numpy().view(np.uint8), bitorder="little")[:n]
            if ref is None:
                ref = got
            print(json.dumps({"env": env, "out": mode, "planes_kept": keep_planes, "ms_min": min(times), "ms_med": sorted(times)[len(times) // 2],
                              "checks_per_s": n / (min(times) * 1e-3), "same_as_first": bool(np.array_equal(ref, got)), "free": float(got.mean())}), flush=True)
    ctx.close()
