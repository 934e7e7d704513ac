#!/usr/bin/env python
"""Where the end-to-end step of the C4 sweep goes: upload, sweep and download timed alone and together (wall clock)."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
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
dev_out = cover_b200.Result(free_bits=torch.empty((n + 31) // 32, dtype=torch.int32, device="cuda"), not_ok=torch.zeros(1, dtype=torch.int64, device="cuda"))


def timed(name, fn, reps=8):
    for _ in range(2):
        fn()
    ctx.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    ctx.synchronize()
    print(json.dumps({"what": name, "ms": (time.perf_counter() - t0) / reps * 1e3}), flush=True)


def both_async():
    cmap.upload_async(pinned)
    cmap.area_is_free(wl.RECT, lat, out=host_out)


def both_sync():
    cmap.upload(pinned)
    cmap.area_is_free(wl.RECT, lat, out=host_out)


def async_dev():
    cmap.upload_async(pinned)
    cmap.area_is_free(wl.RECT, lat, out=dev_out)
    ctx.synchronize()


def upload_async_only():
    cmap.upload_async(pinned)
    ctx.synchronize()


timed("upload (synchronous)", lambda: cmap.upload(pinned))
timed("upload_async + synchronize", upload_async_only)
timed("sweep, device bits", lambda: (cmap.area_is_free(wl.RECT, lat, out=dev_out), ctx.synchronize()))
timed("sweep, host bits", lambda: cmap.area_is_free(wl.RECT, lat, out=host_out))
timed("upload + sweep, host bits", both_sync)
timed("upload_async + sweep, device bits", async_dev)
timed("upload_async + sweep, host bits", both_async)

# host-side cost of one call: time until the call returns (launches queued) vs until the results exist
import cProfile, pstats
for _ in range(3):
    ctx.synchronize()
    t0 = time.perf_counter()
    cmap.area_is_free(wl.RECT, lat, out=dev_out)
    t1 = time.perf_counter()
    ctx.synchronize()
    t2 = time.perf_counter()
    print(json.dumps({"what": "sweep, device bits: call returns / results ready", "ms": [(t1 - t0) * 1e3, (t2 - t0) * 1e3]}), flush=True)
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    cmap.area_is_free(wl.RECT, lat, out=dev_out)
    ctx.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)

# which outputs cost what (device-resident, CUDA events on the library stream)
stream = torch.cuda.ExternalStream(ctx.stream)
variants = {
    "is_free bytes": cover_b200.Result(is_free=torch.empty(n, dtype=torch.uint8, device="cuda")),
    "is_free bytes + status": cover_b200.Result(is_free=torch.empty(n, dtype=torch.uint8, device="cuda"), status=torch.empty(n, dtype=torch.uint8, device="cuda")),
    "free_bits": cover_b200.Result(free_bits=torch.empty((n + 31) // 32, dtype=torch.int32, device="cuda")),
    "free_bits + not_ok": dev_out,
}
for name, o in variants.items():
    for _ in range(2):
        cmap.area_is_free(wl.RECT, lat, out=o)
    ctx.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(stream)
    for _ in range(5):
        cmap.area_is_free(wl.RECT, lat, out=o)
    e.record(stream)
    ctx.synchronize()
    print(json.dumps({"what": "device ms/sweep with outputs: " + name, "ms": s.elapsed_time(e) / 5}), flush=True)
