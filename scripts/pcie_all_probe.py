"""Aggregate pinned H2D / D2H rates with every rank copying at once (torchrun, one rank per GPU)."""
import os, time, torch, torch.distributed as dist
local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
h = torch.empty(16_000_000, dtype=torch.uint8).pin_memory(); d = torch.empty(16_000_000, dtype=torch.uint8, device="cuda")
h2 = torch.empty(12_531_600, dtype=torch.uint8).pin_memory(); d2 = torch.empty(12_531_600, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(name, fn, nbytes, reps=100):
    fn(); torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if rank == 0: print(f"{name}: {world} ranks at once: {world * nbytes * reps / dt / 1e9:.1f} GB/s aggregate, {dt / reps * 1e3:.3f} ms per round", flush=True)
run("H2D 16 MB", lambda: d.copy_(h, non_blocking=True), 16e6)
run("D2H 12.5 MB", lambda: h2.copy_(d2, non_blocking=True), 12.5316e6)
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
run("H2D 16 MB || D2H 12.5 MB", both, 28.53e6)
t = torch.empty(16_000_000, dtype=torch.uint8, device="cuda")
def bc():
    dist.broadcast(t, 0)
run("NCCL broadcast 16 MB", bc, 16e6 / world)
dist.destroy_process_group()
