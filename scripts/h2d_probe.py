import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
import cover_b200
ctx = cover_b200.Context(0)
data = np.zeros((4000, 4000), dtype=np.uint8)
pinned = torch.from_numpy(data).pin_memory()
cmap = cover_b200.Costmap(ctx, data, 0.05)
arr = pinned.numpy()
for name, a in (("pinned", arr), ("pageable", data)):
    for _ in range(3): cmap.upload(a)
    t = time.perf_counter()
    for _ in range(20): cmap.upload(a)
    dt = (time.perf_counter() - t) / 20
    print(name, "upload 16MB ms", dt * 1e3, "GB/s", 16e6 / dt / 1e9)
d = torch.empty(16_000_000, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(20): d.copy_(pinned.view(-1), non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t) / 20
print("torch pinned 1D copy ms", dt * 1e3, "GB/s", 16e6 / dt / 1e9)
h = torch.empty(12_531_600, dtype=torch.uint8).pin_memory()
dd = torch.empty(12_531_600, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(20): h.copy_(dd, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 20
print("torch D2H 12.5MB ms", dt * 1e3, "GB/s", 12.5e6 / dt / 1e9)
