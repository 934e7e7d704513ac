"""Does splitting a 16 MB pinned host-to-device copy over two streams (two copy engines) beat one copy?"""
import time
import torch

n = 16_000_000
src = torch.empty(n, dtype=torch.uint8).pin_memory()
dst = torch.empty(n, dtype=torch.uint8, device="cuda")
streams = [torch.cuda.Stream() for _ in range(4)]


def run(parts, reps=50):
    bounds = [n * k // parts for k in range(parts + 1)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        for k in range(parts):
            with torch.cuda.stream(streams[k]):
                dst[bounds[k]:bounds[k + 1]].copy_(src[bounds[k]:bounds[k + 1]], non_blocking=True)
        torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    print(f"{parts} stream(s): {dt * 1e3:.3f} ms per 16 MB = {n / dt / 1e9:.1f} GB/s", flush=True)


for p in (1, 2, 4, 1, 2):
    run(p)

# both directions at once: 16 MB in while 12.5 MB go out (what two planning cycles in flight ask of the link)
m = 12_531_600
out_dev = torch.empty(m, dtype=torch.uint8, device="cuda")
out_host = torch.empty(m, dtype=torch.uint8).pin_memory()
s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
for label, do_in, do_out in (("H2D only", True, False), ("D2H only", False, True), ("both", True, True), ("both", True, True)):
    torch.cuda.synchronize()
    reps = 50
    t0 = time.perf_counter()
    for _ in range(reps):
        if do_in:
            with torch.cuda.stream(s_in):
                dst.copy_(src, non_blocking=True)
        if do_out:
            with torch.cuda.stream(s_out):
                out_host.copy_(out_dev, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    print(f"{label}: {dt * 1e3:.3f} ms per round", flush=True)
